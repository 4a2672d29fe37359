"""Top stall instructions of one kernel from `ncu -i rep --page source --csv` (instruction, samples, share, stall reasons)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
isrc, iall = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[iall]) for r in data) or 1
print(f"SASS line, instruction, samples, share of all warp samples, top stall reasons   ({sys.argv[2] if len(sys.argv) > 2 else ''})")
for n, r in enumerate(data):
    sm = int(r[iall])
    if sm >= 0.005 * tot:
        top = sorted(((hdr[i], int(r[i] or 0)) for i in stall_cols), key=lambda x: -x[1])[:2]
        print(f"{n},{r[isrc].strip()},{sm},{100 * sm / tot:.1f}%,{[(a, str(b)) for a, b in top if b]}")
