mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_multigpu.py -x -q -k "2gpu and fused" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_fused_2gpu_b.json 2> gpurun_out/bench_fused_2gpu_b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fused_2gpu_b.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), round(d['T_eff_per_gpu'],1), d['gpu_launches'])
PY
