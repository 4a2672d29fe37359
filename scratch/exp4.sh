timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cy in 16 32 64 128; do for w in diffusion2d stokes2d; do
  export CHMY_CY=$cy
  python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/exp4_${w}_c${cy}.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp4_${w}_c${cy}.json')); print('$w CY=$cy', round(d['ms_per_step'],4), round(d['value'],1), d['roofline']['step_kernels_ms'])"
done; done
unset CHMY_CY
python bench.py --workload stokes3d_thermal --steps 15 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stokes3d_thermal', round(d['ms_per_step'],3), round(d['value'],1), d['roofline']['step_kernels_ms'])"
python bench.py --steps 30 --warmup 5 | tee gpurun_out/bench_v4.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stokes3d', round(d['ms_per_step'],3), round(d['value'],1), d['roofline']['step_kernels_ms'], d['e2e']['value'])"
