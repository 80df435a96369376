#!/bin/bash
# upfirdn2d iteration loop: op parity tests, standalone HBM rates, one ncu --set full capture of the three bulk kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -6
timeout 200 python tools/ops_hbm.py 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bulk_kernel --launch-skip 3 -c 3 \
   -o gpurun_out/upfir_bulk python tools/upfir_once.py > gpurun_out/ncu_upfir.log 2>&1; echo "ncu rc=$?"
