#!/bin/bash
# Round 2, call 50: two-operation exact division in the fused 3D sweep (selected per launch when proven for all four divisors)
mkdir -p gpurun_out
set +e
timeout -k 5 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c50_gpu_tests.log
for v in 0 1; do
  echo "== CHMY_DIV2=$v"
  CHMY_DIV2=$v GEOMS='6,4,64,1;6,4,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | grep -v unfused
done | tee gpurun_out/c50_tune_div2.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/c50_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c50_bench.json 2> gpurun_out/c50_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c50_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4), "division:", d["division"], "e2e", round(d["e2e"]["value"], 1))
PY
