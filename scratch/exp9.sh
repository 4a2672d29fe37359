export GEOMS="8,2,64,0;8,2,64,1;16,2,64,0;16,2,64,1;8,4,64,1;8,8,64,1;4,8,64,1;4,4,64,1;16,4,64,1;8,4,128,1"
timeout 600 python scratch/tune_fused.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -3
