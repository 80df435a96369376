#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export HAIRFAST_TEST_DTYPES=default
timeout 300 python -m pytest tests/test_gpu_encoders.py -m gpu -x -q -k "stem3x3 or bisenet_glue" 2>&1 | tail -2
unset HAIRFAST_TEST_DTYPES
timeout 300 python tools/time_stems.py 2>&1 | tail -5
