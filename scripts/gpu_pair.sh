#!/bin/bash
O=gpurun_out
SH="262144,384,1536,0,2,0 262144,384,384,0,2,0 262144,1152,384,0,0,1 313600,1152,384,0,0,1 1048576,192,768,0,2,0 1048576,768,192,1,0,1 1048576,576,192,0,0,1 65536,768,3072,0,2,0 262144,256,256,0,2,0"
echo "== pairs for BN>=256 (default)" > $O/s17_pair.log; python scripts/gemm_probe.py $SH >> $O/s17_pair.log 2>&1
echo "== pairs for BN>=192" >> $O/s17_pair.log; CVB_PAIR_MINBN=192 python scripts/gemm_probe.py $SH >> $O/s17_pair.log 2>&1
echo "== pairs for BN>=128" >> $O/s17_pair.log; CVB_PAIR_MINBN=128 python scripts/gemm_probe.py $SH >> $O/s17_pair.log 2>&1
cat $O/s17_pair.log
