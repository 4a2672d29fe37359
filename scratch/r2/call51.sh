#!/bin/bash
# Round 2, call 51: two-operation division in the 2D stress+velocity sweep too; whole suite + smoke + default bench at HEAD
mkdir -p gpurun_out
set +e
timeout -k 5 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c51_gpu_tests.log
for v in 0 1; do
  echo "== CHMY_DIV2=$v"
  CHMY_DIV2=$v timeout -k 5 200 python scratch/tune_pairs.py stokes2d 2>&1 | grep -E "two kernels|cy=32 "
done | tee gpurun_out/c51_tune_div2_2d.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/c51_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c51_bench.json 2> gpurun_out/c51_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c51_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4), "division:", d["division"], "e2e", round(d["e2e"]["value"], 1), "cpu", round(d["cpu_baseline"]["value"], 1))
for w in d["extra"]["workloads"]:
    print("  ", w.get("workload", "")[:40], "fused", w["fused"], round(w.get("ms_per_step", 0), 3), "ms", round(w.get("T_eff", 0), 1), "GB/s", round(w.get("frac_of_hbm_peak", 0), 3), w.get("error", ""))
PY
