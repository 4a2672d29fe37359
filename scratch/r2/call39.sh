#!/bin/bash
# Round 2, call 39: 4-operation exact division (double-double reciprocal product + one Markstein correction)
mkdir -p gpurun_out
set +e
timeout -k 5 600 python -m pytest tests/test_b200_parity.py tests/test_b200_fused.py tests/test_b200_fused2d.py tests/test_golden_fixtures.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c39_tests.log
GEOMS='6,4,64,1;6,4,64,1;4,6,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c39_tune.log
