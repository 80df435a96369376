#!/bin/bash
# GPU session: the full HairFast.swap() arms (stock reference / overlay / overlay_fast) and the swap test.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
t0=$(date +%s)
timeout 1500 python -m pytest tests/test_gpu_swap.py -m gpu -x -q -s > gpurun_out/pytest_swap.log 2>&1; echo "pytest swap rc=$? ($(( $(date +%s) - t0 )) s)"
tail -40 gpurun_out/pytest_swap.log
cat gpurun_out/swap_arms_*.json 2>/dev/null | head -150
tail -3 gpurun_out/parity.jsonl 2>/dev/null
