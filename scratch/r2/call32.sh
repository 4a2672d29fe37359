#!/bin/bash
# Round 2, call 32: tau_old (and tau) of the next plane requested after the stresses of a plane, before the barrier
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_b200_fused.py -m gpu -q -x -k "ahead or geom" 2>&1 | tail -4 | tee gpurun_out/c32_fused_tests.log
GEOMS='6,4,64,1;6,4,64,9;6,4,64,17;4,6,64,9;4,6,64,17;6,3,64,9;6,5,64,9;6,4,32,9;6,4,128,9;6,4,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | grep -v unfused | tee gpurun_out/c32_tune_ahead.log
