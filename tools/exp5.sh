#!/bin/bash
for cfg in "HF_MMA_MODE=1 HF_CONV_2CTA=0" "HF_MMA_MODE=2 HF_CONV_2CTA=0"; do
  echo "=== $cfg"
  env $cfg timeout 300 python tools/diag_conv.py 2>&1 | grep -v "^   " | grep -c "max_err/rms=0.01"
  env $cfg timeout 200 python tools/prof_chain.py 4 2> gpurun_out/prof_x.txt; tail -29 gpurun_out/prof_x.txt | grep -E "conv +#(6|7|8|9|10|11|12|13|14|15|16) |sum of"
done
