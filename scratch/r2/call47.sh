#!/bin/bash
# Round 2, call 47: final check after the thermal sweep's occupancy change: whole suite, smoke, default bench
mkdir -p gpurun_out
set +e
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c47_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/c47_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c47_bench.json 2> gpurun_out/c47_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c47_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4), "launches/step", d["launches_per_step"], "traffic", d["roofline"]["traffic"])
print("e2e", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("what", "steady")})
for w in d["extra"]["workloads"]:
    print("  ", w.get("workload", "")[:40], "fused", w["fused"], round(w.get("ms_per_step", 0), 3), "ms", round(w.get("T_eff", 0), 1), "GB/s", round(w.get("frac_of_hbm_peak", 0), 3), w.get("error", ""))
PY
