#!/bin/bash
# Are the hot kernels still the ones a B200 measured?  Compiles ops_fast.cu, ops_fused.cu and ops_fast2d.cu of the last
# GPU-run commit (default c8c3df1, round 1 session 3) and of HEAD with the library's flags and compares the SASS of every
# kernel present in both, instruction by instruction (addresses and encodings stripped).  Needs nvcc only (no GPU).
#   bash scratch/sass_vs_measured.sh [commit]
set -e
BASE=${1:-c8c3df1}
W=$(mktemp -d)
mkdir -p $W/old $W/new
git archive $BASE chmy.jl_b200/csrc include | tar -x -C $W/old
git archive HEAD chmy.jl_b200/csrc include | tar -x -C $W/new
for v in old new; do
  for f in ops_fast ops_fused ops_fast2d ops; do
    ( cd $W/$v/chmy.jl_b200/csrc
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true \
          -prec-sqrt=true -ccbin /usr/bin/g++ -cubin -o $W/$v/$f.cubin $f.cu
      cuobjdump -sass $W/$v/$f.cubin | grep -E "Function :|^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > $W/$v/$f.sass ) &
  done
done
wait
python3 - $W <<'PY'
import re, sys
w = sys.argv[1]
def split(path):
    d, cur = {}, None
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []
        elif cur:
            d[cur].append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line).strip())
    return d
bad = 0
for f in ("ops_fast", "ops_fused", "ops_fast2d", "ops"):
    o, n = split(f"{w}/old/{f}.sass"), split(f"{w}/new/{f}.sass")
    diff = [k for k in o if k in n and o[k] != n[k]]
    gone = [k for k in o if k not in n]
    print(f"{f}: {len([k for k in o if k in n and o[k] == n[k]])} kernels identical, {len(diff)} different, "
          f"{len(gone)} removed, {len([k for k in n if k not in o])} new")
    for k in diff + gone:
        print("   ", k)
    bad += len(diff) + len(gone)
sys.exit(1 if bad else 0)
PY
rm -rf $W
