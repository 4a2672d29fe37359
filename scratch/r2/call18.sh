#!/bin/bash
mkdir -p gpurun_out
set +e
timeout 600 python -m pytest tests/test_b200_fused.py -q -x -k "any_geometry" 2>&1 | tail -3
CHMY_FUSE_VARIANT=3 CHMY_FUSE_TYB=4 CHMY_FUSE_CL=4 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scratch/run_fused_once.py 70 37 20 1 2>&1 | tail -4
GEOMS='6,4,64,1;4,4,64,3;4,6,64,3;6,4,64,3;6,2,64,3;4,3,64,3;4,6,128,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c18_tune_fused.log
CHMY_FUSE_VARIANT=3 CHMY_FUSE_TYB=4 CHMY_FUSE_CL=6 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o gpurun_out/c18_tma_full \
    python scratch/run_fused_once.py 767 767 255 3 > gpurun_out/c18_tma_full.log 2>&1
ncu -i gpurun_out/c18_tma_full.ncu-rep --page raw --csv > gpurun_out/c18_tma_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c18_tma_full_raw.csv | tee gpurun_out/c18_tma_summary.csv | head -30
