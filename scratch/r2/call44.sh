#!/bin/bash
# Round 2, call 44 (2 GPUs): interiors of the last T layers of tiles run last (T = 1, 2, 3): what is left exposed of the exchange?
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_b200_fused.py tests/test_split_plan.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/c44_tests.log
timeout 300 python -m pytest tests/test_z_b200_multigpu.py -q -x -k "2gpu and fused" 2>&1 | tail -3 | tee -a gpurun_out/c44_tests.log
for T in 1 2 3 2; do
  CHMY_TAIL_LAYERS=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 30 --warmup 5 --split on --no-e2e --no-check > gpurun_out/c44_bench_2gpu_T$T.json 2> gpurun_out/c44_bench_2gpu_T$T.err
  python - "$T" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c44_bench_2gpu_T{sys.argv[1]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("T", sys.argv[1], round(d["ms_per_step"], 3), "ms/iter", "overlapped", d["overlapped_launches"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex); print(open(f"gpurun_out/c44_bench_2gpu_T{sys.argv[1]}.err").read()[-1500:])
PY
done | tee gpurun_out/c44_tail_layers.log
CHMY_TAIL_LAYERS=2 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('1 GPU', round(d['ms_per_step'],3))" | tee -a gpurun_out/c44_tail_layers.log
