#!/bin/bash
# Lean GPU session: changed-area tests (default operand types only), op HBM rates, one bench line without the
# child-process comparator legs, the overlay swap arm, and the ncu captures of the hi-res generator kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HAIRFAST_TEST_DTYPES=default
t0=$(date +%s)
timeout 900 python -m pytest ${LEAN_TESTS:-tests/test_gpu_ops.py tests/test_glue.py tests/test_gpu_encoders.py tests/test_gpu_conv.py tests/test_gpu_generator.py} -m gpu -x -q > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -15 gpurun_out/pytest_lean.log
unset HAIRFAST_TEST_DTYPES
timeout 300 python tools/ops_hbm.py 2>&1 | tail -4
t0=$(date +%s)
timeout 900 python bench.py --no-comparators --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_lean.json 2> gpurun_out/bench_lean.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/bench_lean.err; cat gpurun_out/bench_lean.json
if [ "${LEAN_SWAP:-1}" = "1" ]; then
  t0=$(date +%s)
  timeout 900 python baseline/run_swap.py --mode overlay --work /tmp/hairfast_work --reps 5 --warmup 3 > gpurun_out/swap_overlay.json 2> gpurun_out/swap_overlay.err; echo "swap rc=$? ($(( $(date +%s) - t0 )) s)"
  tail -2 gpurun_out/swap_overlay.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/swap_overlay.json') if l.startswith('{')][-1])
for t in d['timings']: print('wall %.1f gpu %.1f hot %.1f'%(t['wall_ms'],t['gpu_ms'],t['hot_path_ms']), {k:round(v,2) for k,v in t['per_module_ms'].items()})
"
fi
if [ "${LEAN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo_kernel -c 8 -f \
      -o gpurun_out/hires_r2 python tools/ncu_gen.py 4 > gpurun_out/ncu_hires.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_hires.log
fi
