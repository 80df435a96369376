#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
