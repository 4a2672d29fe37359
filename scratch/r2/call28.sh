#!/bin/bash
# Round 2, call 28: lagged march with the small shared-memory footprint (2-slot ring + 7-entry stash)
mkdir -p gpurun_out
set +e
timeout 900 python -m pytest tests/test_b200_fused.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c28_fused_tests_lag.log
for c in default 58; do
  echo "== carveout $c %"
  if [ $c = default ]; then unset CHMY_FUSE_CARVEOUT; else export CHMY_FUSE_CARVEOUT=$c; fi
  CHMY_DEBUG_OCC=1 GEOMS='6,4,64,1;6,4,64,3;4,6,64,3;4,4,64,3;6,3,64,3;6,5,64,3;6,4,32,3;6,4,128,3;6,4,64,1' timeout 600 python scratch/tune_fused.py 2>&1 | grep -v unfused
done | tee gpurun_out/c28_tune_lag.log
