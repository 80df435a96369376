#!/bin/bash
# ncu evidence of the round: full-set capture of the dominant kernel at the step's batch (B = 96), and the launch list
# (gpu__time_duration) of exactly one bench step at T = 4.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo --launch-skip 5 -c 1 -f \
    -o gpurun_out/dominant_b96 python tools/ncu_dominant.py 96 > gpurun_out/ncu_dominant_b96.log 2>&1; echo "ncu dominant rc=$?"; tail -2 gpurun_out/ncu_dominant_b96.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step_T16.csv python bench.py --profile-step --triples 16 > gpurun_out/ncu_step.log 2>&1
echo "ncu launch list rc=$? lines=$(wc -l < gpurun_out/launches_step_T16.csv)"
