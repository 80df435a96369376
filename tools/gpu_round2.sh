#!/bin/bash
# One GPU session of round 2: full -m gpu suite (both operand types + the swap arms), smoke, bench with all legs.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
t0=$(date +%s)
timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
