for ty in 8 16; do for cz in 12 16 24 32 48; do
  export CHMY_TY=$ty CHMY_CZ=$cz CHMY_SYNC=1
  python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/exp3_t${ty}_c${cz}.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp3_t${ty}_c${cz}.json')); print('TY=$ty CZ=$cz', round(d['ms_per_step'],3), d['roofline']['step_kernels_ms'])"
done; done
