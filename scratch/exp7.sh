mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_b200_fused.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_b200_parity.py -x -q -k "valued_bc or bc_" 2>&1 | tail -5
timeout 600 python scratch/tune_fused.py 2>&1 | tail -30
export CHMY_FUSE_TYB=8 CHMY_FUSE_CL=2 CHMY_FUSE_CZ=64 CHMY_FUSE_PF=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_sv -s 2 -c 1 -f -o gpurun_out/fused_511_t8c2pf python scratch/run_fused_once.py 511 511 511 3 2>&1 | tail -3
