export GEOMS="8,2,64,1;4,8,64,1;4,4,64,1;4,2,64,1;4,1,64,1;4,8,128,1"
timeout 600 python scratch/tune_fused.py 2>&1 | tail -12
