mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -5
export CHMY_FUSE_TYB=8 CHMY_FUSE_CL=2 CHMY_FUSE_CZ=64
python scratch/run_fused_once.py 511 511 511
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/fused_511_t8c2 python scratch/run_fused_once.py 511 511 511 3 2>&1 | tail -5
export CHMY_FUSE_TYB=16 CHMY_FUSE_CL=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/fused_511_t16c2 python scratch/run_fused_once.py 511 511 511 3 2>&1 | tail -5
ls -la gpurun_out
