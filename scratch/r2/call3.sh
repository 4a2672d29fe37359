#!/bin/bash
# Round 2, GPU call 3: first run of the TMA-fed sweep (variant bit 1): parity, memcheck, A/B.
mkdir -p gpurun_out
set +e
CHMY_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_b200_fused.py -q -x -k experimental 2>&1 | tail -15 | tee gpurun_out/c3_fused_tests.log
CHMY_FUSE_VARIANT=3 CHMY_FUSE_TYB=4 CHMY_FUSE_CL=4 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scratch/run_fused_once.py 70 37 20 1 2>&1 | tail -15 | tee gpurun_out/c3_memcheck.log
GEOMS='6,4,64,1;4,4,64,3;4,6,64,3;6,4,64,3;4,3,64,3;4,2,64,3;6,2,64,3;4,4,128,3;4,4,32,3;8,4,64,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | tee gpurun_out/c3_tune_fused.log
