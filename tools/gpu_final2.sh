#!/bin/bash
# Last session of the round (short budget): whole -m gpu suite + smoke + standalone operator rates.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
t0=$(date +%s)
timeout 420 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -9 gpurun_out/pytest_gpu.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 60 python tools/ops_hbm.py 2>&1 | tail -3
