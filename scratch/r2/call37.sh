#!/bin/bash
# Round 2, call 37: phase A's loads issued in consumption order (volatile loads)
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_b200_fused.py -m gpu -q -x -k "any_geometry" 2>&1 | tail -3 | tee gpurun_out/c37_tests.log
GEOMS='6,4,64,1;6,4,64,3;4,6,64,3;6,4,64,1;6,4,64,3' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | grep -v unfused | tee gpurun_out/c37_tune_ordered.log
