#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/diag_conv.py 2>&1 | grep -v "^   " | tail -14
echo "=== chain profile v2"; timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_v2.txt; tail -29 gpurun_out/prof_v2.txt
echo "=== chain profile v1"; HF_CONV_V1=1 timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_v1.txt; tail -29 gpurun_out/prof_v1.txt | grep -v rgb
