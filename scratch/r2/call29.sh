#!/bin/bash
# Round 2, call 29: ncu --set full of the lagged march (why is it slow?)
mkdir -p gpurun_out
set +e
export CHMY_FUSE_VARIANT=3
( time timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/c29_lag_full \
    python scratch/run_fused_once.py 767 767 255 2 > gpurun_out/c29_full.log 2>&1 ) 2>&1 | tail -3
ncu -i gpurun_out/c29_lag_full.ncu-rep --page raw --csv > gpurun_out/c29_lag_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c29_lag_full_raw.csv | tee gpurun_out/c29_lag_summary.csv | head -40
ls -la gpurun_out/c29_lag_full.ncu-rep
