#!/bin/bash
# One full GPU session of round 2: the whole -m gpu suite (both operand types + the swap arms) with durations, smoke,
# the reference arm, bench with all legs.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -22 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$? ($(( $(date +%s) - t0 )) s)"
cut -c1-1500 gpurun_out/bench_reference.json
t0=$(date +%s)
timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
