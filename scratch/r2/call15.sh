#!/bin/bash
# Round 2, GPU call 15: suite on HEAD + the ncu evidence of the shipped sweep (6 rows x cluster 4, slim halo rows) at 767^3.
mkdir -p gpurun_out
set +e
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c15_gpu_tests.log
echo "== dram bytes of one sweep at 767^3"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none \
    -k regex:k_fused_sv -s 2 -c 2 --csv --log-file gpurun_out/c15_fused_767_dram_bytes.csv python scratch/run_fused_once.py 767 767 767 2 > gpurun_out/c15_ncu_run.log 2>&1
tail -6 gpurun_out/c15_fused_767_dram_bytes.csv | cut -c1-260
echo "== launch list of the bench command"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c15_launches_bench_767.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/c15_launches.log 2>&1
python scratch/ncu_summary.py launches gpurun_out/c15_launches_bench_767.csv | tee gpurun_out/c15_launches_summary.txt
echo "== ncu --set full of one sweep at 767^3"
( time timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/c15_fused_767_full \
    python scratch/run_fused_once.py 767 767 767 2 > gpurun_out/c15_fused_full.log 2>&1 ) 2>&1 | tail -3
ncu -i gpurun_out/c15_fused_767_full.ncu-rep --page raw --csv > gpurun_out/c15_fused_767_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c15_fused_767_full_raw.csv | tee gpurun_out/c15_fused_767_summary.csv | head -40
ls -la gpurun_out/c15_fused_767_full.ncu-rep
