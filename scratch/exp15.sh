mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_b200_multigpu.py -x -q -k "8gpu and fused" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_fused_8gpu.json 2> gpurun_out/bench_fused_8gpu.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_fused_8gpu.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], round(d['ms_per_step'],3), round(d['T_eff_per_gpu'],1), d['gpu_launches'], d['config']['proc_dims'])
except Exception as e:
    print("no line", e); print(open('gpurun_out/bench_fused_8gpu.err').read()[-2000:])
PY
