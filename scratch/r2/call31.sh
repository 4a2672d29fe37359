#!/bin/bash
# Round 2, call 31: neighbour handshake through mbarriers (no L1 invalidation per plane), plain and lagged march
mkdir -p gpurun_out
set +e
timeout -k 5 200 python -m pytest tests/test_b200_fused.py -m gpu -q -x -k "mbar" 2>&1 | tail -4 | tee gpurun_out/c31_fused_tests.log
timeout -k 5 300 python -m pytest tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/c31_fused_tests.log
GEOMS='6,4,64,1;6,4,64,5;6,4,64,3;6,4,64,7;4,6,64,5;4,6,64,7;6,3,64,5;6,5,64,5;6,6,64,5;6,8,64,5;6,4,64,1' timeout -k 5 150 python scratch/tune_fused.py 2>&1 | grep -v unfused | tee gpurun_out/c31_tune_mbar.log
