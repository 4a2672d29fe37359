#!/bin/bash
# Round 2, multi-GPU call (N = 2 first, then 8):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash scratch/r2_multigpu_call.sh 2'
# A/B of the launch order on connected ranks: overlapped inner + slabs (the reference's order, 94.7 % weak-scaling
# efficiency in round 1), one full-range sweep followed by the batches / exchange, and the library's self-tuning default.
N=${1:-2}
mkdir -p gpurun_out
set +e
for mode in auto on off; do
  flag="--split $mode"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-e2e $flag > gpurun_out/r2_bench_${N}gpu_${mode}.json 2> gpurun_out/r2_bench_${N}gpu_${mode}.err
  python - "$N" "$mode" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "GPUs", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d["config"]["proc_dims"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e)
    print(open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu_{sys.argv[2]}.err").read()[-1500:])
PY
done
CHMY_SPLIT=0 timeout 400 python -m pytest tests/test_z_b200_multigpu.py -x -q -k "${N}gpu" 2>&1 | tail -4

# EXPERIMENTAL peer-store transport (comm.cu, CHMY_EXCHANGE_PEER; protocol proven by tests/test_peer_protocol.py): first GPU
# run.  Waits give up after 10 s (CHMY_PEER_TIMEOUT_S) and every command has its own timeout, so a wrong hand-shake cannot
# hang the box.  Parity first (gated cases), then the A/B against NCCL with the same launch order.
export CHMY_PEER_TIMEOUT_S=10
CHMY_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_z_b200_multigpu.py -x -q -k "${N}gpu and peer" 2>&1 | tail -6
for xm in nccl peer; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --split on --exchange $xm > gpurun_out/r2_bench_${N}gpu_x${xm}.json 2> gpurun_out/r2_bench_${N}gpu_x${xm}.err
  python - "$N" "$xm" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu_x{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "GPUs exchange", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d.get("exchange_msgs"))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no line:", e)
    print(open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu_x{sys.argv[2]}.err").read()[-1500:])
PY
done
