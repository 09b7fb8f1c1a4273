#!/bin/bash
O=gpurun_out
SH="4194304,384,96,1,0,1 1048576,768,192,1,0,1 262144,1536,384,1,0,1 65536,3072,768,1,0,1"
echo "== sigmoid-form GELU (2 MUFU)" > $O/s9_tanh.log; python scripts/gemm_probe.py $SH >> $O/s9_tanh.log 2>&1
echo "== tanh-form GELU (1 MUFU)" >> $O/s9_tanh.log; CVB_GELU_TANH=1 python scripts/gemm_probe.py $SH >> $O/s9_tanh.log 2>&1
echo "== diag sig" >> $O/s9_tanh.log; python scripts/diag_sam2.py tiny 3 2>&1 | grep -E "^low|^masks|^image|^trunk3" >> $O/s9_tanh.log
echo "== diag tanh" >> $O/s9_tanh.log; CVB_GELU_TANH=1 python scripts/diag_sam2.py tiny 3 2>&1 | grep -E "^low|^masks|^image|^trunk3" >> $O/s9_tanh.log
cat $O/s9_tanh.log
