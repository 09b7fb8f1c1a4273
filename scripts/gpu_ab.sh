#!/bin/bash
# same-box A/B of one env switch on the SAM 2.1 forward: usage gpu_ab.sh ENVVAR [tag]
O=gpurun_out
V=$1; T=${2:-ab}
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_maps_gpu.py tests/test_sam2_golden_gpu.py tests/test_sam2_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/${T}_pytest.log
for i in 1 2; do
  env $V=1 python scripts/profile_sam2.py tiny 64 64 > $O/${T}_on$i.log 2>&1
  env $V=0 python scripts/profile_sam2.py tiny 64 64 > $O/${T}_off$i.log 2>&1
done
env $V=1 python scripts/profile_sam2.py base_plus 64 64 > $O/${T}_bp_on.log 2>&1
env $V=0 python scripts/profile_sam2.py base_plus 64 64 > $O/${T}_bp_off.log 2>&1
tail -2 $O/${T}_pytest.log
head -1 $O/${T}_on1.log $O/${T}_off1.log $O/${T}_on2.log $O/${T}_off2.log $O/${T}_bp_on.log $O/${T}_bp_off.log
