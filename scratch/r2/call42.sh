#!/bin/bash
# Round 2, call 42 (2 GPUs): multi-GPU parity incl. peer-store cases (now always on), bench with per-dimension exchange timing and the fallback counter.
N=${1:-2}
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests/test_z_b200_multigpu.py -q -x -k "${N}gpu" 2>&1 | tail -8 | tee gpurun_out/c42_multigpu_tests_${N}.log
for sp in on off; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --split $sp > gpurun_out/c42_bench_${N}gpu_${sp}.json 2> gpurun_out/c42_bench_${N}gpu_${sp}.err
  python - "$N" "$sp" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c42_bench_{sys.argv[1]}gpu_{sys.argv[2]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "GPUs split", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d["config"]["proc_dims"],
          "launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"], "fallbacks", d.get("fusion_fallbacks"), "e2e", round(d["e2e"]["value"], 1))
    print("   exchange:", d.get("exchange_alone"))
    print("   check:", d["multi_gpu_check"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e)
    print(open(f"gpurun_out/c42_bench_{sys.argv[1]}gpu_{sys.argv[2]}.err").read()[-2500:])
PY
done
