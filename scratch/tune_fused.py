"""In-process sweep over the fused sweep's tile geometry at the headline size (scratch tool, not a bench line)."""
import math, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chmy_b200 as ch
from chmy_b200 import drivers as BD

n = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (767, 767, 767)
arch = ch.Arch(ch.B200Backend())
sol = BD.Stokes(arch, n, re_m=2.5 * math.pi, rho_g_function=True, outer_width=(128, 8, 4), adv_coef=0.01, blocking=False)
sol.begin_time_step()
A = 37 * 8.0 * n[0] * n[1] * n[2]

def timeit(K=8, W=2):
    for _ in range(W):
        sol.mechanics()
    ch.synchronize(arch)
    ch.event_record(arch, 0)
    for _ in range(K):
        sol.mechanics()
    ch.event_record(arch, 1)
    ch.synchronize(arch)
    return ch.event_elapsed_ms(arch, 0, 1) / K

ch.set_fusion(arch, False)
ms = timeit()
print(f"unfused            : {ms:8.3f} ms  T_eff {A/ms/1e6:8.1f} GB/s", flush=True)
ch.set_fusion(arch, True)
geoms = [tuple(int(x) for x in g.split(',')) for g in os.environ.get('GEOMS', '8,2,64,0;8,2,64,1').split(';')]
out = {}
for g in geoms:
    ch.set_fused_tuning(arch, *g)
    try:
        ms = timeit()
    except Exception as e:
        print(g, "FAILED", e, flush=True)
        continue
    out[str(g)] = ms
    print(f"fused tyb,cl,cz,pf={g!s:14}: {ms:8.3f} ms  T_eff {A/ms/1e6:8.1f} GB/s", flush=True)
json.dump(out, open("gpurun_out/tune_fused.json", "w"))
arch.close()
