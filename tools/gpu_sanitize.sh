#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (upfirdn2d cp.async rings, fused stems, label arg-max,
# F-space glue, bicubic / dilate) + smoke.  One dtype (the default) to bound the run; logs under gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
CS="compute-sanitizer --tool memcheck --print-limit 20"
run() {  # name, pytest args...
    local name=$1; shift
    local t0=$(date +%s)
    timeout 700 $CS python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_$name.log 2>&1
    echo "$name rc=$? ($(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$name.log | tr '\n' ' ')"
}
run ops tests/test_gpu_ops.py -k "not full_size"
run glue tests/test_glue.py
run seg_stems tests/test_gpu_encoders.py -k "glue_kernels or fused_stem3x3 or bisenet_golden"
t0=$(date +%s)
timeout 400 $CS python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.log 2>&1
echo "smoke rc=$? ($(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_smoke.log)"
