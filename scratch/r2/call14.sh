#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
set +e
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv | head -10
for tag in on1 on2 off1; do
  sp=${tag%?}
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-check --split $sp > gpurun_out/c14_bench_${N}gpu_${tag}.json 2> gpurun_out/c14_bench_${N}gpu_${tag}.err
  python - "$N" "$tag" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c14_bench_{sys.argv[1]}gpu_{sys.argv[2]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "GPUs", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter by rank", [round(x, 3) for x in d["ms_per_step_by_rank"]], "sub", d["roofline"]["step_kernels_ms"], "xchg", d["exchange_alone"].get("ms_per_exchange_alone"))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e)
    print(open(f"gpurun_out/c14_bench_{sys.argv[1]}gpu_{sys.argv[2]}.err").read()[-2500:])
PY
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv | head -10
