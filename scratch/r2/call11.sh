#!/bin/bash
# 8-GPU call: parity of the 4- and 8-rank cases, bench on/off (on with e2e)
mkdir -p gpurun_out
set +e
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/c11_topo.txt
nproc > gpurun_out/c11_nproc.txt; free -g | head -2 >> gpurun_out/c11_nproc.txt; numactl -H 2>/dev/null | head -8 >> gpurun_out/c11_nproc.txt
CHMY_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_z_b200_multigpu.py -q -x -k "4gpu or 8gpu" 2>&1 | tail -6 | tee gpurun_out/c11_multigpu_tests.log
for N in 8 4; do
for sp in on off; do
  extra="--no-e2e"; if [ "$sp" = "on" ] && [ "$N" = "8" ]; then extra=""; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 $extra --split $sp > gpurun_out/c11_bench_${N}gpu_${sp}.json 2> gpurun_out/c11_bench_${N}gpu_${sp}.err
  python - "$N" "$sp" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/c11_bench_{sys.argv[1]}gpu_{sys.argv[2]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "GPUs split", sys.argv[2], round(d["ms_per_step"], 3), "ms/iter", round(d["T_eff_per_gpu"], 1), "GB/s/GPU", d["config"]["proc_dims"],
          "launches/step", d["launches_per_step"], "overlapped", d["overlapped_launches"], "check:", d["multi_gpu_check"]["ok"], d["multi_gpu_check"].get("max_rel"),
          "e2e", {k: round(e.get(k, 0), 1) for k in ("value", "upload_ms", "iterate_ms", "download_ms")})
except Exception as ex:
    print(sys.argv[1], sys.argv[2], "no line:", ex)
    print(open(f"gpurun_out/c11_bench_{sys.argv[1]}gpu_{sys.argv[2]}.err").read()[-2500:])
PY
done
done
