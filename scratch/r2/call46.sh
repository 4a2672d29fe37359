#!/bin/bash
# Round 2, call 46: 2D stress+velocity sweep at 4 / 5 / 6 resident CTAs per SM
mkdir -p gpurun_out
set +e
for occ in 4 5 6; do
  echo "== occ $occ"
  CHMY_FUSE2D_OCC=$occ timeout -k 5 200 python scratch/tune_pairs.py stokes2d 2>&1 | grep -E "two kernels|cy=16 |cy=32 |cy=64 "
done | tee gpurun_out/c46_sv2_occupancy.log
