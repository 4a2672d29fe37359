#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
set +e
timeout 600 python -m pytest tests/test_b200_fused.py -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_z_b200_multigpu.py -q -x -k "${N}gpu and stokes_fused" 2>&1 | tail -3
for sp in on off; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --split $sp > gpurun_out/c10_bench_${N}gpu_${sp}.json 2> gpurun_out/c10_bench_${N}gpu_${sp}.err
  python - "$N" "$sp" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c10_bench_{sys.argv[1]}gpu_{sys.argv[2]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "GPUs split", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d["config"]["proc_dims"],
          "launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"], "check ok:", d["multi_gpu_check"]["ok"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e)
    print(open(f"gpurun_out/c10_bench_{sys.argv[1]}gpu_{sys.argv[2]}.err").read()[-2500:])
PY
done
