#!/bin/bash
# Round 2, call 35: staged uploads (contiguous pieces + scatter kernel): parity of the copy paths, bandwidth, e2e bench
mkdir -p gpurun_out
set +e
timeout -k 5 300 python -m pytest tests/test_zz_b200_round2.py tests/test_b200_parity.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/c35_tests.log
timeout -k 5 200 python scratch/copy_bw.py 2>&1 | tail -7 | tee gpurun_out/c35_copy_bandwidth.log
CHMY_NO_STAGED_UPLOAD=1 timeout -k 5 200 python scratch/copy_bw.py 2>&1 | grep "set!" | sed 's/^/strided path: /' | tee -a gpurun_out/c35_copy_bandwidth.log
timeout -k 5 420 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c35_bench.json 2> gpurun_out/c35_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c35_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
e = d["e2e"]
print("headline", round(d["ms_per_step"], 3), "ms; e2e", round(e["value"], 1), "GB/s; upload", round(e["upload_ms"], 1), "iterate", round(e["iterate_ms"], 1), "download", round(e["download_ms"], 1))
PY
