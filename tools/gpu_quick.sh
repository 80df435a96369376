#!/bin/bash
# Quick kernel iteration: conv / generator / encoder parity tests, per-launch chain timing at B=4, bench lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py tests/test_gpu_encoders.py -m gpu -x -q 2>&1 | tail -3
unset HAIRFAST_TEST_DTYPES
HF_GEN_PROFILE=1 HAIRFAST_CUDA_GRAPHS=0 timeout 300 python tools/prof_chain.py 4 2>&1 | grep "forward 2" -A40 | grep -E "conv  |sum of" | awk '{print $3,$4,$6,$7,$11,$12}' | tr '\n' ';'
echo
for T in ${QUICK_T:-16}; do
timeout 600 python bench.py --no-comparators --no-cpu-baseline --no-extras --triples $T 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('T=$T value',d['value'],'e2e',d['e2e']['value'],'T1',d['latency_T1']['latency_ms_per_triple'],'clk',d['clocks'])"
done
