#!/bin/bash
# memcheck + racecheck over the op tests after the TMA bulk-copy upfirdn2d kernels went in.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
for tool in memcheck racecheck; do
  t0=$(date +%s)
  timeout 100 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_${tool}_ops_bulk.log 2>&1
  echo "$tool rc=$? ($(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_ops_bulk.log | tr '\n' ' ')"
done
