#!/bin/bash
# ncu --set full of the current GEMM on representative encoder shapes and of one global-attention launch
O=gpurun_out
REPS=1 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc -c 12 -o $O/s41_gemm python scripts/gemm_probe.py 4194304,96,384,0,2,0 4194304,384,96,1,0,1 1048576,768,192,1,0,1 262144,1536,384,1,0,1 262144,384,1536,0,2,0 65536,3072,768,1,0,1 > $O/s41_gemm_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_attn_global -c 1 -o $O/s41_attn python scripts/profile_sam2.py tiny 16 16 > $O/s41_attn_ncu.log 2>&1
tail -2 $O/s41_gemm_ncu.log; tail -2 $O/s41_attn_ncu.log
