#!/bin/bash
# Round 2, call 43 (8 GPUs): the final state -- parity of the 8-rank cases, bench overlapped (with e2e) and one-stream
mkdir -p gpurun_out
set +e
timeout 400 python -m pytest tests/test_z_b200_multigpu.py -q -x -k "8gpu" 2>&1 | tail -4 | tee gpurun_out/c43_multigpu_tests.log
for sp in on off; do
  extra="--no-e2e"; if [ "$sp" = "on" ]; then extra=""; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup 5 $extra --split $sp > gpurun_out/c43_bench_8gpu_${sp}.json 2> gpurun_out/c43_bench_8gpu_${sp}.err
  python - "$sp" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c43_bench_8gpu_{sys.argv[1]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    e = d.get("e2e") or {}
    print("8 GPUs split", sys.argv[1], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d["config"]["proc_dims"],
          "launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"], "check:", d["multi_gpu_check"]["ok"], d["multi_gpu_check"].get("max_rel"),
          "e2e", {k: round(e.get(k, 0), 1) for k in ("value", "upload_ms", "iterate_ms", "download_ms")}, "exchange", (d.get("exchange_alone") or {}).get("ms_per_exchange_alone"), (d.get("exchange_alone") or {}).get("ms_per_dim_alone"))
except Exception as ex:
    print(sys.argv[1], "no line:", ex)
    print(open(f"gpurun_out/c43_bench_8gpu_{sys.argv[1]}.err").read()[-2500:])
PY
done
