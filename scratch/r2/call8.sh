#!/bin/bash
# Round 2, GPU call 8: overlap (boundary tiles first + retire counter), folded batches, frame-copy elision, per-ctx tuning.
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -25 | tee gpurun_out/c8_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/c8_smoke.log
for sp in on off; do
  timeout 420 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --split $sp > gpurun_out/c8_bench_$sp.json 2> gpurun_out/c8_bench_$sp.err
  python - $sp <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c8_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("split", sys.argv[1], round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"], d["roofline"]["step_kernels_ms"])
except Exception as e:
    print("no line:", e); print(open(f"gpurun_out/c8_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c8_launches_stokes3d_767.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c8_launches.log 2>&1
python scratch/ncu_summary.py launches gpurun_out/c8_launches_stokes3d_767.csv | tee gpurun_out/c8_launches_summary.txt
