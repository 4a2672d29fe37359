#!/bin/bash
mkdir -p gpurun_out
set +e
for sp in on always on always; do
  timeout 420 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --split $sp > gpurun_out/c21_bench_$sp.json 2> gpurun_out/c21_bench_$sp.err
  python - $sp <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c21_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("split", sys.argv[1], round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"])
except Exception as e:
    print("no line:", e); print(open(f"gpurun_out/c21_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
