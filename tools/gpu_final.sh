#!/bin/bash
# Final session of the round: whole -m gpu suite, smoke, bench with all legs (default flags), ncu launch list of a step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
t0=$(date +%s)
timeout 1700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
t0=$(date +%s)
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step_T16.csv python bench.py --profile-step --triples 16 > gpurun_out/ncu_step.log 2>&1
echo "ncu launch list rc=$? lines=$(wc -l < gpurun_out/launches_step_T16.csv)"
