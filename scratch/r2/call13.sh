#!/bin/bash
mkdir -p gpurun_out
set +e
timeout 600 python -m pytest tests/test_b200_fused.py -q -x 2>&1 | tail -2
GEOMS='6,4,64,1;6,4,64,5;4,6,64,5;6,5,64,5;6,4,64,1;6,4,64,5' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c13_tune_fused.log
