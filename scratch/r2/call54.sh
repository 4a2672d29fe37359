#!/bin/bash
# Round 2, call 54: bench with the sweep kernel timed on its own (events inside the library around the kernel)
mkdir -p gpurun_out
set +e
timeout -k 5 120 python -m pytest tests/test_b200_fused.py -m gpu -q -x -k "fused_iteration and geom3" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/c54_bench.json 2> gpurun_out/c54_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c54_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
r = d["roofline"]
print("headline", round(d["ms_per_step"], 3), "ms; roofline frac", round(r["frac"], 4), "kernel_ms", round(r["kernel_ms"], 3), {k[:40]: round(v, 3) for k, v in r["step_kernels_ms"].items()}, "share", round(r["share_of_step"], 3))
PY
tail -3 gpurun_out/c54_bench.err
