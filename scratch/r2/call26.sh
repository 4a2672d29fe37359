#!/bin/bash
mkdir -p gpurun_out
set +e
CHMY_DEBUG_OCC=1 GEOMS='6,4,64,1;6,4,64,3;4,6,64,3;4,6,64,1' timeout 600 python scratch/tune_fused.py 767 767 255 2>&1 | tee gpurun_out/c26_tune_lag.log
