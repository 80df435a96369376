#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
HF_SWAP_TRACE=2 HAIRFAST_CUDA_GRAPHS=1 timeout 900 python baseline/run_swap.py --mode overlay --work /tmp/hairfast_work --reps 12 --warmup 2 2>&1 >/dev/null | grep TRACE | cut -c1-230
