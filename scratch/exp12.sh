mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_multigpu.py -x -q -k "2gpu" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_fused_2gpu.json 2> gpurun_out/bench_fused_2gpu.err; tail -c 1500 gpurun_out/bench_fused_2gpu.json; tail -3 gpurun_out/bench_fused_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --fused 0 --no-e2e > gpurun_out/bench_unfused_2gpu.json 2> gpurun_out/bench_unfused_2gpu.err; tail -c 600 gpurun_out/bench_unfused_2gpu.json
