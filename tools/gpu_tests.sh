#!/bin/bash
# The whole -m gpu suite as the driver runs it (both operand types + the swap arms), with durations; then op rates.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
t0=$(date +%s)
timeout 1700 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/ops_hbm.py 2>&1 | tail -4
