mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -3
export GEOMS="4,4,64,1;4,4,96,1;4,8,64,1;4,2,64,1;8,2,64,1"
timeout 600 python scratch/tune_fused.py 2>&1 | tail -7
