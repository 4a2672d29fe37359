#!/bin/bash
# Round 2, call 36: k_bc_all with 8 face points per thread
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_b200_parity.py tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c36_tests.log
timeout -k 5 300 python bench.py --steps 30 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c36_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s", d["roofline"]["kernel_ms"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c36_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
python scratch/ncu_summary.py launches gpurun_out/c36_launches.csv 2>/dev/null | head -5
