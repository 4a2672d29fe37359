#!/bin/bash
# Round 2, call 40: warm-up / closing planes of a z-chunk load only what is consumed
mkdir -p gpurun_out
set +e
timeout -k 5 600 python -m pytest tests/test_b200_fused.py tests/test_zy_b200_fullsize.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c40_tests.log
GEOMS='6,4,64,1;6,4,32,1;6,4,48,1;6,4,96,1;6,4,128,1;4,6,64,1;6,4,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c40_tune.log
