#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/ops_hbm.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upfirdn2d --launch-skip 3 -c 1 -f -o gpurun_out/upfirdn_up1 python tools/ops_hbm.py > gpurun_out/ncu_ops.log 2>&1; echo rc=$?
