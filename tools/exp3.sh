#!/bin/bash
mkdir -p gpurun_out
echo "=== v3 pitch=${HF_HALO_PITCH:-10}"; timeout 300 python tools/diag_conv.py 2>&1 | grep -v "^   " | tail -19
echo "=== chain profile v3"; timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_v3.txt; tail -29 gpurun_out/prof_v3.txt
if grep -q "max_err/rms=[1-9]" gpurun_out/diag.txt; then
  echo "=== pitch 16 fallback"; HF_HALO_PITCH=16 timeout 300 python tools/diag_conv.py 2>&1 | grep -v "^   " | tail -19
fi
