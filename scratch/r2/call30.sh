#!/bin/bash
# Round 2, call 30: lagged march, tau loads before the velocity update and V loads after it
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c30_fused_tests_lag.log
CHMY_DEBUG_OCC=1 GEOMS='6,4,64,1;6,4,64,3;4,6,64,3;4,4,64,3;6,3,64,3;6,5,64,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | grep -v unfused | tee gpurun_out/c30_tune_lag.log
export CHMY_FUSE_VARIANT=3
timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/c30_lag_full \
    python scratch/run_fused_once.py 767 767 255 2 > gpurun_out/c30_full.log 2>&1
ncu -i gpurun_out/c30_lag_full.ncu-rep --page raw --csv > gpurun_out/c30_lag_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c30_lag_full_raw.csv | tee gpurun_out/c30_lag_summary.csv | grep -i "duration\|dram\|stalled\|hit"
