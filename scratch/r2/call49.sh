#!/bin/bash
# Round 2, call 49: ncu --set full of the final fused sweep and of the thermal sweep at 767^3
mkdir -p gpurun_out
set +e
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/c49_fused_767_full \
    python scratch/run_fused_once.py 767 767 767 2 > gpurun_out/c49_full.log 2>&1
ncu -i gpurun_out/c49_fused_767_full.ncu-rep --page raw --csv > gpurun_out/c49_fused_767_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c49_fused_767_full_raw.csv | tee gpurun_out/c49_fused_767_summary.csv | head -30
ncu -i gpurun_out/c49_fused_767_full.ncu-rep --page source --csv > gpurun_out/c49_src.csv 2>/dev/null
python scratch/top_stalls.py gpurun_out/c49_src.csv "ncu --set full --import-source on, final k_fused_sv<0,1,6> at 767^3" | tee gpurun_out/c49_top_stalls.txt | head -20
rm -f gpurun_out/c49_src.csv
