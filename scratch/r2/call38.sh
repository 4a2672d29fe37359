#!/bin/bash
# Round 2, call 38: TIMING EXPERIMENT -- how much of the sweep is the exact-division sequence? (variants are not exact; never shipped)
mkdir -p gpurun_out
set +e
cp chmy.jl_b200/libchmy_b200.so /tmp/lib_ship.so
for v in ship div1 div2; do
  if [ $v = ship ]; then cp /tmp/lib_ship.so chmy.jl_b200/libchmy_b200.so; else cp scratch/libchmy_b200_$v.so chmy.jl_b200/libchmy_b200.so; fi
  touch chmy.jl_b200/libchmy_b200.so
  echo "== $v"
  GEOMS='6,4,64,1;6,4,64,1' timeout -k 5 200 python scratch/tune_fused.py 2>&1 | grep -v "^unfused" 
done | tee gpurun_out/c38_division_cost.log
cp /tmp/lib_ship.so chmy.jl_b200/libchmy_b200.so
