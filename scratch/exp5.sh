mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -15
timeout 600 python scratch/tune_fused.py 2>&1 | tail -30
