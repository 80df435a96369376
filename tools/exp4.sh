#!/bin/bash
for cfg in "HF_HALO_NA=2 HF_HALO_STAGES=4" "HF_HALO_NA=2" "HF_HALO_STAGES=4" "HF_HALO_GMAX=1"; do
  echo "=== $cfg"
  env $cfg timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_x.txt; tail -29 gpurun_out/prof_x.txt | grep -E "conv +#(4|6|8|10|12|14|16) |sum of"
done
