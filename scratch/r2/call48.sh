#!/bin/bash
# Round 2, call 48 (8 GPUs): interiors of the last T layers of tiles last (T = 1, 2, 3) with three connected dimensions
mkdir -p gpurun_out
set +e
for T in 1 2 1 2; do
  CHMY_TAIL_LAYERS=$T timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 30 --warmup 5 --split on --no-e2e --no-check > gpurun_out/c48_bench_8gpu_T$T.json 2> gpurun_out/c48_bench_8gpu_T$T.err
  python - "$T" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c48_bench_8gpu_T{sys.argv[1]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("T", sys.argv[1], round(d["ms_per_step"], 3), "ms/iter", "overlapped", d["overlapped_launches"], d.get("ms_per_step_by_rank"))
except Exception as ex:
    print(sys.argv[1], "no line:", ex); print(open(f"gpurun_out/c48_bench_8gpu_T{sys.argv[1]}.err").read()[-1500:])
PY
done | tee gpurun_out/c48_tail_layers_8gpu.log
