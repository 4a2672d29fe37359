export GEOMS="8,2,64,0;8,2,64,1;8,1,64,0;16,1,64,0;8,4,128,0;4,4,64,0"
timeout 600 python scratch/tune_fused.py 2>&1 | tail -12
