"""A few PT iterations at a given size with the fused sweep (for ncu captures; tuning via CHMY_FUSE_* env)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chmy_b200 as ch
from chmy_b200 import drivers as BD
n = tuple(int(x) for x in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
arch = ch.Arch(ch.B200Backend())
ch.set_fusion(arch, os.environ.get("FUSED", "1") == "1")
sol = BD.Stokes(arch, n, re_m=2.5 * math.pi, rho_g_function=True, outer_width=(128, 8, 4), adv_coef=0.01, blocking=False)
sol.begin_time_step()
for _ in range(iters):
    sol.mechanics()
ch.synchronize(arch)
ch.event_record(arch, 0)
for _ in range(iters):
    sol.mechanics()
ch.event_record(arch, 1)
ch.synchronize(arch)
print("ms/iter", ch.event_elapsed_ms(arch, 0, 1) / iters, "fused sweeps", ch.fused_count(arch))
arch.close()
