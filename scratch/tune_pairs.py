"""In-process sweeps over the tuning knobs of the EXPERIMENTAL flux->update sweeps (scratch tool for round 2, not a bench line).

    python scratch/tune_pairs.py stokes2d|diffusion2d|stokes2d_thermal|stokes3d_thermal [n...]

2D sweeps: rows per y-chunk x rows per load group (chmy_set_fused2d_tuning); 3D thermal sweep: planes per z-chunk
(chmy_set_fused2d_tuning).  Prints ms per PT iteration next to the two-kernel time of the same process.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chmy_b200 as ch
from chmy_b200 import drivers as BD

wl = sys.argv[1] if len(sys.argv) > 1 else "stokes2d"
default_n = {"stokes2d": (8191, 8191), "stokes2d_thermal": (8191, 8191), "diffusion2d": (16383, 16383), "stokes3d_thermal": (767, 767, 767)}[wl]
n = tuple(int(x) for x in sys.argv[2:]) or default_n
nio = {"stokes2d": 22, "stokes2d_thermal": 29, "diffusion2d": 7, "stokes3d_thermal": 46}[wl]
A = nio * 8.0 * float(np.prod(n))
arch = ch.Arch(ch.B200Backend())
if wl == "diffusion2d":
    sol = BD.Diffusion2D(arch, n, outer_width=(128, 8), C0=None, blocking=False)
    rng = np.random.default_rng(0)
    for j0 in range(1, n[1] + 1, 2048):
        j1 = min(n[1], j0 + 2047)
        sol.C.from_host(rng.random((n[0], j1 - j0 + 1)), [1, j0], [n[0], j1])
    step = sol.step
else:
    sol = BD.Stokes(arch, n, re_m=2.5 * math.pi, rho_g_function=True, outer_width=(128, 8, 4)[:len(n)], adv_coef=0.01, blocking=False)
    sol.begin_time_step()
    if wl.endswith("_thermal"):
        def step():
            sol.mechanics()
            sol.thermal()
    else:
        step = sol.mechanics


def timeit(K=10, W=3):
    for _ in range(W):
        step()
    ch.synchronize(arch)
    ch.event_record(arch, 0)
    for _ in range(K):
        step()
    ch.event_record(arch, 1)
    ch.synchronize(arch)
    return ch.event_elapsed_ms(arch, 0, 1) / K


def show(tag, ms):
    print(f"{tag:34}: {ms:8.3f} ms  T_eff {A / ms / 1e6:8.1f} GB/s", flush=True)


ch.set_fusion(arch, 1 if len(n) == 3 else 0)
show("two kernels per pair", timeit())
ch.set_fusion(arch, 3)
if len(n) == 2:
    for cy in (16, 32, 64, 128, 256):
        for un in ((1, 2, 4) if wl != "stokes2d" else (1,)):
            ch.set_fused2d_tuning(arch, cy, un)
            try:
                show(f"fused cy={cy} unroll={un}", timeit())
            except Exception as e:
                print("FAILED", cy, un, e, flush=True)
else:
    for cz in (4, 8, 12, 16, 24, 32, 64):
        ch.set_fused2d_tuning(arch, 0, 0, cz)
        try:
            show(f"fused thermal sweep cz={cz}", timeit())
        except Exception as e:
            print("FAILED", cz, e, flush=True)
print("fused sweeps launched:", ch.fused_count(arch))
arch.close()
