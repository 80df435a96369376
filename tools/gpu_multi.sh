#!/bin/bash
# N-GPU session: the 2-GPU output-equality test and one torchrun bench line (as the driver launches it).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L | head -8
export HAIRFAST_TEST_DTYPES=default
timeout 900 python -m pytest tests/test_bench_contract.py -m gpu -q -k two_gpu 2>&1 | tail -3
unset HAIRFAST_TEST_DTYPES
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','nccl_broadcast_bytes_at_init','shard_output_equality','clocks')}, d['e2e'])"
