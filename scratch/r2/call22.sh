#!/bin/bash
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 420 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c22_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4))
for w in d["extra"]["workloads"]:
    print("  ", w.get("workload", "")[:40], "fused", w["fused"], round(w.get("ms_per_step", 0), 3), "ms", round(w.get("T_eff", 0), 1), "GB/s", w.get("error", ""))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c22_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
python scratch/ncu_summary.py launches gpurun_out/c22_launches.csv | head -5
