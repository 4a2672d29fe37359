#!/bin/bash
mkdir -p gpurun_out
set +e
for wl in stokes3d_thermal stokes2d diffusion2d stokes2d_thermal; do
  echo "== $wl"
  timeout 400 python scratch/tune_pairs.py $wl 2>&1 | tee gpurun_out/c20_tune_pairs_$wl.log
done
