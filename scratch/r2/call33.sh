#!/bin/bash
# Round 2, call 33: L2 bulk prefetch of the next plane's operands (no registers, no scoreboards)
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_b200_fused.py -m gpu -q -x -k "l2 or geom" 2>&1 | tail -4 | tee gpurun_out/c33_fused_tests.log
GEOMS='6,4,64,1;6,4,64,9;6,4,64,17;4,6,64,9;4,6,64,17;6,3,64,9;6,5,64,9;6,4,32,9;6,4,128,9;6,4,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | grep -v unfused | tee gpurun_out/c33_tune_l2.log
