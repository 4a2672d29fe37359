#!/bin/bash
mkdir -p gpurun_out
set +e
timeout 600 python -m pytest tests/test_b200_parity.py tests/test_zz_b200_round2.py tests/test_zy_b200_fullsize.py -q -x -k "roundtrip or pinned or fullsize or 767 or set_" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c19_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"])
print("e2e", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("what", "steady")})
PY
