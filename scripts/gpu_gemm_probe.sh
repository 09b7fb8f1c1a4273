#!/bin/bash
O=gpurun_out
SH="4194304,96,96,0,2,0 4194304,384,96,1,0,1 4194304,576,96,0,0,1 4194304,96,384,0,2,0 1048576,768,192,1,0,1 1048576,192,192,0,2,0 262144,1536,384,1,0,1 262144,384,1536,0,2,0 262144,384,384,0,2,0"
python scripts/gemm_probe.py $SH > $O/s6_gemm_plain.log 2>&1
REPS=1 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc -c 8 -o $O/s6_gemm python scripts/gemm_probe.py 4194304,96,96,0,2,0 4194304,384,96,1,0,1 1048576,768,192,1,0,1 262144,1536,384,1,0,1 > $O/s6_gemm_ncu.log 2>&1
cat $O/s6_gemm_plain.log
