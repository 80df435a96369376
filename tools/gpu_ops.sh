#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/ops_hbm.py 2>&1 | tail -4
