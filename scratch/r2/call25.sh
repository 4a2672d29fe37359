#!/bin/bash
# Round 2, call 25: lagged march with a split barrier that has work on both sides -- parity on the GPU, timing at 767^3.
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c25_fused_tests_lag.log
GEOMS='6,4,64,1;6,4,64,3;4,6,64,3;4,4,64,3;6,3,64,3;6,5,64,3;6,4,32,3;6,4,128,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c25_tune_lag.log
