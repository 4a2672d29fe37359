#!/bin/bash
mkdir -p gpurun_out
set +e
for c in default 38 44 50 58 65 72 86 100; do
  echo "== carveout $c %"
  if [ $c = default ]; then unset CHMY_FUSE_CARVEOUT; else export CHMY_FUSE_CARVEOUT=$c; fi
  CHMY_DEBUG_OCC=1 GEOMS='6,4,64,1;6,4,64,3' timeout 600 python scratch/tune_fused.py 767 767 255 2>&1 | grep -v unfused
done | tee gpurun_out/c27_carveout.log
