#!/bin/bash
mkdir -p gpurun_out
set +e
CHMY_FUSE_VARIANT=3 CHMY_FUSE_TYB=4 CHMY_FUSE_CL=4 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o gpurun_out/c4_tma_full \
    python scratch/run_fused_once.py 767 767 255 3 > gpurun_out/c4_tma_full.log 2>&1
ncu -i gpurun_out/c4_tma_full.ncu-rep --page raw --csv > gpurun_out/c4_tma_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c4_tma_full_raw.csv | tee gpurun_out/c4_tma_summary.csv | head -40
