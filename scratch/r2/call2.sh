#!/bin/bash
# Round 2, GPU call 2: cluster sizes and cache-policy flavours of the fused sweep (A/B only, not a bench line).
mkdir -p gpurun_out
set +e
CHMY_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_b200_fused.py -q -x 2>&1 | tail -5 | tee gpurun_out/c2_fused_tests.log
GEOMS='6,4,64,1;6,3,64,1;6,5,64,1;6,6,64,1;6,8,64,1;4,6,64,1;8,4,64,1;8,3,64,1;6,4,64,5;6,4,64,9;6,4,64,13;6,4,64,25;6,4,64,29;4,4,64,5;4,4,64,13;4,4,64,29;6,4,128,1;6,4,32,1;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c2_tune_fused.log
