mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -3
export GEOMS="4,4,64,1;4,4,64,0;4,4,48,1;4,4,96,1;4,4,32,1;4,8,64,1;4,2,64,1"
timeout 600 python scratch/tune_fused.py 2>&1 | tail -9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/fused_511_t4c4 python scratch/run_fused_once.py 511 511 511 3 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_fused_v1.json 2> gpurun_out/bench_fused_v1.err; tail -c 3000 gpurun_out/bench_fused_v1.json; tail -5 gpurun_out/bench_fused_v1.err
