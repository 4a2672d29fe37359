#!/bin/bash
# Round 2, call 24: the lagged march (velocity of plane kp-2 under the loads of plane kp) -- parity on the GPU, then timing at 767^3.
mkdir -p gpurun_out
set +e
CHMY_FUSE_VARIANT=3 timeout 600 python -m pytest tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c24_fused_tests_lag.log
GEOMS='6,4,64,1;6,4,64,3;6,4,64,7;4,6,64,3;4,4,64,3;6,3,64,3;6,5,64,3;6,4,32,3;6,4,128,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c24_tune_lag.log
