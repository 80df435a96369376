#!/bin/bash
# GPU experiment: validate the halo kernel (shifted UMMA descriptors) and time the chain per launch.
mkdir -p gpurun_out
echo "=== halo, base_offset=1 (documented)"; HF_HALO_BASE_OFFSET=1 timeout 200 python tools/diag_conv.py 2>&1 | grep -v "^   " | tail -16
cp gpurun_out/diag.txt gpurun_out/diag_bo1.txt
echo "=== halo, base_offset=0"; HF_HALO_BASE_OFFSET=0 timeout 200 python tools/diag_conv.py 2>&1 | grep -v "^   " | tail -16
cp gpurun_out/diag.txt gpurun_out/diag_bo0.txt
echo "=== chain profile v1"; HF_CONV_V1=1 timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_v1.txt; tail -34 gpurun_out/prof_v1.txt
echo "=== chain profile v2 bo=${BO:-1}"; HF_HALO_BASE_OFFSET=${BO:-1} timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_v2.txt; tail -34 gpurun_out/prof_v2.txt
