#!/bin/bash
O=gpurun_out
SH="4194304,96,96,0,2,0 4194304,96,168,0,2,0 1048576,192,192,0,2,0 4194304,96,384,0,2,0"
echo "== 8 epilogue warps" > $O/s22_ew.log; python scripts/gemm_probe.py $SH >> $O/s22_ew.log 2>&1
echo "== 16 epilogue warps (K<=192, f32+res)" >> $O/s22_ew.log; CVB_GEMM_EW16_RES=1 python scripts/gemm_probe.py $SH >> $O/s22_ew.log 2>&1
CVB_GEMM_EW16_RES=1 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $O/s22_ew.log
cat $O/s22_ew.log
