#!/bin/bash
# Round 2, call 53: ncu launch list of the bench command at HEAD
mkdir -p gpurun_out
set +e
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c53_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
python scratch/ncu_summary.py launches gpurun_out/c53_launches.csv 2>/dev/null | head -6 | tee gpurun_out/c53_launches_summary.txt
