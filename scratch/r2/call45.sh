#!/bin/bash
# Round 2, call 45: 3D thermal sweep at 2 / 3 / 4 resident CTAs per SM
mkdir -p gpurun_out
set +e
for occ in 2 3 4; do
  echo "== occ $occ"
  CHMY_FUSE_T3_OCC=$occ timeout -k 5 200 python scratch/tune_pairs.py stokes3d_thermal 2>&1 | grep -E "two kernels|cz=32|cz=64"
done | tee gpurun_out/c45_t3_occupancy.log
CHMY_FUSE_T3_OCC=3 timeout -k 5 200 python -m pytest tests/test_b200_fused2d.py -m gpu -q -x -k "thermal or t3" 2>&1 | tail -2
