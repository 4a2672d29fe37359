#!/bin/bash
# Round 2, GPU call 1 (one B200): everything written after round 1's GPU budget ran out + the A/B of the sweep candidates.
mkdir -p gpurun_out
set +e
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/c1_gpu.txt
echo "== 1. whole -m gpu suite, experimental cases included"
CHMY_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -rfEs --durations=8 2>&1 | tail -60 | tee gpurun_out/c1_gpu_tests.log

echo "== 2. bench line (default workload)"
timeout 420 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_stokes3d.json 2> gpurun_out/c1_bench_stokes3d.err
tail -c 1800 gpurun_out/c1_bench_stokes3d.json; tail -5 gpurun_out/c1_bench_stokes3d.err

echo "== 3. A/B of the fused-sweep candidates at 767^3 (rows, cluster, z-chunk, variant bit0 relaxed arrive / bit1 pipelined)"
GEOMS='4,4,64,1;6,2,64,1;12,1,64,1;6,4,64,1;2,8,64,1;2,4,64,1;4,4,64,3;4,2,64,3;4,1,64,3;2,8,64,3;2,4,64,3;4,4,128,1;4,4,192,1;4,4,64,1' timeout 500 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c1_tune_fused.log

echo "== 4. 2D workloads and 3D thermal: two kernels vs the experimental sweeps"
for wl in stokes2d diffusion2d stokes2d_thermal; do
  for fu in 0 3; do
    timeout 200 python bench.py --workload $wl --fused $fu --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/c1_${wl}_f${fu}.json 2> gpurun_out/c1_${wl}_f${fu}.err
    python - "$wl" "$fu" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c1_{sys.argv[1]}_f{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "fused" if sys.argv[2] != "0" else "two-kernel", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s", d["roofline"]["step_kernels_ms"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e); print(open(f"gpurun_out/c1_{sys.argv[1]}_f{sys.argv[2]}.err").read()[-800:])
PY
  done
done
for fu in 1 3; do
  timeout 300 python bench.py --workload stokes3d_thermal --fused $fu --steps 20 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/c1_stokes3d_thermal_f${fu}.json 2> gpurun_out/c1_stokes3d_thermal_f${fu}.err
  python - "$fu" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c1_stokes3d_thermal_f{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("stokes3d_thermal --fused", sys.argv[1], round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s", d["roofline"]["step_kernels_ms"])
except Exception as e:
    print("stokes3d_thermal", sys.argv[1], "no line:", e); print(open(f"gpurun_out/c1_stokes3d_thermal_f{sys.argv[1]}.err").read()[-800:])
PY
done

echo "== 5. dram bytes of ONE fused sweep at 767^3 (roofline.traffic)"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none \
    -k regex:k_fused_sv -s 2 -c 2 --csv --log-file gpurun_out/c1_fused_767_dram_bytes.csv python scratch/run_fused_once.py 767 767 767 2 > gpurun_out/c1_ncu_run.log 2>&1
tail -12 gpurun_out/c1_fused_767_dram_bytes.csv

echo "== 6. launch list of the bench command"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c1_launches_stokes3d_767.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_launches.log 2>&1
python scratch/ncu_summary.py launches gpurun_out/c1_launches_stokes3d_767.csv | tee gpurun_out/c1_launches_summary.txt

echo "== 7. ncu --set full of one fused sweep at 767x767x255 (memory small enough for kernel replay)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/c1_fused_full \
    python scratch/run_fused_once.py 767 767 255 3 > gpurun_out/c1_fused_full.log 2>&1
ncu -i gpurun_out/c1_fused_full.ncu-rep --page raw --csv > gpurun_out/c1_fused_full_raw.csv 2>/dev/null
python scratch/ncu_summary.py raw gpurun_out/c1_fused_full_raw.csv | tee gpurun_out/c1_fused_summary.csv | head -70
ls -la gpurun_out | head -40
