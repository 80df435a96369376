#!/bin/bash
# Second sanitizer pass: racecheck (shared-memory hazards) over the smem-ring / staging kernels, memcheck over the network
# goldens (tcgen05 conv kernels + glue through the real forwards) and synccheck on the op tests.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
run() {  # tool, name, pytest args...
    local tool=$1 name=$2; shift 2
    local t0=$(date +%s)
    timeout 500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_${tool}_$name.log 2>&1
    echo "$tool $name rc=$? ($(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_$name.log | tr '\n' ' ')"
}
run racecheck ops tests/test_gpu_ops.py -k "not full_size"
run racecheck seg_stems tests/test_gpu_encoders.py -k "glue_kernels or fused_stem3x3 or bisenet_golden"
run racecheck glue tests/test_glue.py
run synccheck ops tests/test_gpu_ops.py -k "not full_size"
run memcheck ops_full tests/test_gpu_ops.py -k "full_size"
run memcheck generator tests/test_gpu_generator.py -k "generator256 or skip_argument or graph_replay or fse_generator"
run memcheck encoders tests/test_gpu_encoders.py -k "e4e_encoder_golden or fse_encoder_golden or feature_encoder_mult or feature_iresnet or conv2d_building_block"
run memcheck conv tests/test_gpu_conv.py
