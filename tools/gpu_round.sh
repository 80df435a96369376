#!/bin/bash
# One GPU session: parity tests (bf16 + fp16), bench line, ncu launch list of one step, ncu full capture of the
# dominant kernel.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_bf16.log 2>&1; echo "pytest bf16 rc=$?"
tail -3 gpurun_out/pytest_gpu_bf16.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --profile-step --triples 4 > gpurun_out/ncu_step.log 2>&1
echo "ncu launch list rc=$? lines=$(wc -l < gpurun_out/launches_step.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo --launch-skip 5 -c 1 -f \
    -o gpurun_out/dominant python tools/ncu_dominant.py > gpurun_out/ncu_dominant.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_dominant.log
HAIRFAST_DTYPE=fp16 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fp16.log 2>&1; echo "pytest fp16 rc=$?"
tail -3 gpurun_out/pytest_gpu_fp16.log
