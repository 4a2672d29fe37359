for h in 0 1 2 3; do for y in 0 1; do
  export CHMY_HINT=$h CHMY_SYNC=$y
  python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/exp_h${h}_s${y}.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp_h${h}_s${y}.json')); print('HINT=$h SYNC=$y', round(d['ms_per_step'],3), d['roofline']['step_kernels_ms'])"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_stress3|k_velocity3" -s 4 -c 2 --csv --log-file gpurun_out/exp_dram_h${h}_s${y}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  grep -E "k_stress3|k_velocity3" gpurun_out/exp_dram_h${h}_s${y}.csv | awk -F'","' '{print $5, $13, $15}'
done; done
