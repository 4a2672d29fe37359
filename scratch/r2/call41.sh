#!/bin/bash
# Round 2, GPU call 41 (final state of round 2): whole GPU suite (Float32 solver ops, BASELINE parity grids), smoke, default bench with extras, reference arm.
mkdir -p gpurun_out
set +e
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 | tee gpurun_out/c41_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/c41_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c41_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", round(d["ms_per_step"], 3), "ms", round(d["T_eff_per_gpu"], 1), "GB/s frac", round(d["roofline"]["frac"], 4), "launches/step", d["launches_per_step"], "traffic", d["roofline"]["traffic"])
    print("e2e", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("what", "steady")})
    print("cpu_baseline", d["cpu_baseline"])
    for w in d["extra"]["workloads"]:
        print("  ", w.get("workload", "")[:40], "fused", w["fused"], round(w.get("ms_per_step", 0), 3), "ms", round(w.get("T_eff", 0), 1), "GB/s", round(w.get("frac_of_hbm_peak", 0), 3), w.get("error", ""))
except Exception as e:
    print("no line:", e); print(open("gpurun_out/c41_bench.err").read()[-2000:])
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c41_ref.json 2> gpurun_out/c41_ref.err ) 2>&1 | tail -3
tail -c 1200 gpurun_out/c41_ref.json; tail -3 gpurun_out/c41_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c41_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
python scratch/ncu_summary.py launches gpurun_out/c41_launches.csv 2>/dev/null | head -6 | tee gpurun_out/c41_launches_summary.txt
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_fused_sv -s 2 -c 2 --csv --log-file gpurun_out/c41_fused_767_dram_bytes.csv python scratch/run_fused_once.py 767 767 767 2 > /dev/null 2>&1
tail -4 gpurun_out/c41_fused_767_dram_bytes.csv | cut -c1-220
