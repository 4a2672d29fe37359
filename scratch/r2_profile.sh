#!/bin/bash
# Round 2 profiling call (one GPU, ~12 GPU-minutes): the evidence round 1 could not capture for the fused sweep at 767^3.
#   /usr/local/graft/bin/gpurun --timeout 1000 -- 'bash scratch/r2_profile.sh'
# Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
set +e
echo "== launch list of the bench command (kernel shares of a step; cold-cache, serialised)"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_stokes3d_767.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
python scratch/ncu_summary.py launches gpurun_out/r2_launches_stokes3d_767.csv | tee gpurun_out/r2_launches_summary.txt

echo "== ncu --set full of ONE k_fused_sv launch at 767^3 (source-level, bring the .ncu-rep back)"
timeout 480 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/r2_fused_767_full \
    python scratch/run_fused_once.py 767 767 767 3 > gpurun_out/r2_fused_full.log 2>&1
ncu -i gpurun_out/r2_fused_767_full.ncu-rep --page raw --csv > gpurun_out/r2_fused_767_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/r2_fused_767_full_raw.csv | tee gpurun_out/r2_fused_767_summary.csv
