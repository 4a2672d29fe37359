for ty in 8 16; do for cz in 64 100 200; do
  export CHMY_TY=$ty CHMY_CZ=$cz CHMY_SYNC=1
  python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/exp2_t${ty}_c${cz}.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp2_t${ty}_c${cz}.json')); print('TY=$ty CZ=$cz', round(d['ms_per_step'],3), d['roofline']['step_kernels_ms'])"
done; done
export CHMY_TY=16 CHMY_CZ=100
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_stress3|k_velocity3" -s 4 -c 2 --csv --log-file gpurun_out/exp2_dram.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -E "k_stress3|k_velocity3" gpurun_out/exp2_dram.csv | awk -F'","' '{print $5, $13, $15}'
