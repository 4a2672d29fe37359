"""
CPU proof of the fused 3D thermal sweep (chmy.jl_b200/csrc/fused_thermal3.cuh; EXPERIMENTAL, chmy_set_fusion(ctx, 3)).

tests/emul/fused_emul_t3.cpp runs the kernel's own load / compute functions lane by lane over every warp of the launch
grid.  The result must be bit-identical to the oracle's update_thermal_flux! on [0, n+1]^3 followed by update_thermal!
on the box (examples/stokes_3d_inc_ve_T.jl:59-77): qT.x, qT.y, qT.z and the new T on the box, everything else -- the
current T buffer included -- untouched.  This pins the recomputation of the +x / +y / +z neighbour fluxes, the "outside
the op's range the update sees the STORED flux" rule at the upper edges, split-launch sub-boxes, odd edges and z-chunks.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_fused_emulation import Pitched

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "fused_emul_t3.cpp")
LIB = os.path.join(HERE, "emul", "libfused_emul_t3.so")
HDRS = [os.path.join(HERE, "..", "chmy.jl_b200", "csrc", h) for h in ("fused_thermal3.cuh", "fused_pairs2d.cuh", "fused_sv.cuh")]


@pytest.fixture(scope="module")
def emul3():
    from helpers import build_emul
    lib = build_emul("fused_emul_t3")
    lib.fused_emul_t3_run.restype = C.c_int
    return lib


def same(a, b, name):
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    if not ok.all():
        idx = np.argwhere(~ok)
        i = tuple(idx[0])
        raise AssertionError(f"{name}: {len(idx)} cells differ, first at storage {i} (logical {tuple(x - 1 for x in i)}): "
                             f"oracle {a[i]!r} fused {b[i]!r}")


def run_case(o, emul3, n, box, cz, rows, seed=0):
    rng = np.random.default_rng(seed)
    g = o.Grid((-1.0, -1.1, -1.2), (2.0, 2.3, 2.6), n)
    T, To = o.Field(g, 0), o.Field(g, 0)
    q, V = o.VectorField(g), o.VectorField(g)
    named = {"T": T, "To": To, "qx": q["x"], "qy": q["y"], "qz": q["z"], "Vx": V["x"], "Vy": V["y"], "Vz": V["z"]}
    for f in named.values():
        f.data[...] = rng.random(f.sdims) - 0.5              # interior, halo AND padding
    if min(n) >= 4:
        V["x"].data[3, 4, 2] = 0.0
        V["y"].data[5, 2, 3] = -0.0                          # max(v, 0) / min(v, 0) at signed zeros
        V["z"].data[2, 3, 4] = np.nan                        # Julia's max/min propagate NaN
    cur = {k: Pitched(f.data) for k, f in named.items()}
    orig = {k: f.data.copy() for k, f in named.items()}
    new_T = cur["T"].copy()
    lam, dt = 0.7, 0.013
    lo, hi = box
    o.update_thermal_flux(g, (q, T, V, lam), (0, 0, 0), tuple(x + 1 for x in n))
    o.update_thermal(g, (T, To, q, dt), lo, tuple(h - 1 for h in hi))

    ptrs = [cur["T"].p0(), new_T.p0(), cur["To"].p0(), cur["qx"].p0(), cur["qy"].p0(), cur["qz"].p0(),
            cur["Vx"].p0(), cur["Vy"].p0(), cur["Vz"].p0()]
    P = (C.c_void_p * 9)(*ptrs)
    for a, b in (("To", "T"), ("Vz", "T"), ("qz", "T"), ("Vx", "qx"), ("Vy", "qy")):
        assert (cur[a].sy, cur[a].sz) == (cur[b].sy, cur[b].sz)
    strides = (C.c_int * 6)(cur["T"].sy, cur["T"].sz, cur["qx"].sy, cur["qx"].sz, cur["qy"].sy, cur["qy"].sz)
    bx = (C.c_int * 12)(*lo, *hi, 0, 0, 0, *(x + 2 for x in n))
    sc = (C.c_double * 5)(lam, dt, *g.inv_spacing)
    assert emul3.fused_emul_t3_run(P, strides, bx, sc, cz, rows) == 0

    sl = tuple(slice(l + 1, h + 1) for l, h in zip(lo, hi))  # logical -> storage index (+1)
    same(T.data[sl], new_T.view()[sl], "T")
    for c in "xyz":
        # the oracle wrote the fluxes on the whole range; the sweep stores them on its box only
        same(q[c].data[sl], cur["q" + c].view()[sl], "qT." + c)
        a, b = orig["q" + c].copy(), cur["q" + c].view().copy()
        a[sl] = 0.0
        b[sl] = 0.0
        same(a, b, "outside-box qT." + c)
    a, b = orig["T"].copy(), new_T.view().copy()
    a[sl] = 0.0
    b[sl] = 0.0
    same(a, b, "outside-box T (shadow)")
    for k in ("T", "To", "Vx", "Vy", "Vz"):                  # read-only for the kernel (T: the current buffer)
        same(orig[k], cur[k].view(), "read-only " + k)


CASES = [
    # n, box (lo, hi exclusive) or None for the full range, planes per z-chunk, rows per CTA
    ((70, 13, 9), None, 4, 8),
    ((70, 13, 9), None, 64, 8),
    ((125, 21, 6), None, 3, 4),
    ((61, 9, 11), None, 5, 8),
    ((9, 5, 4), None, 2, 8),
    ((1, 1, 1), None, 1, 8),
    ((130, 20, 12), ((8, 4, 3), (124, 18, 11)), 4, 8),         # inner region of a split launch
    ((130, 11, 10), ((0, 0, 0), (132, 13, 3)), 8, 8),          # bottom z slab
    ((130, 11, 10), ((0, 0, 9), (132, 13, 12)), 2, 8),         # top z slab (reads the stored q.z[n+2])
    ((66, 20, 8), ((0, 16, 3), (68, 22, 7)), 16, 8),           # top y slab (stored q.y[n+2])
    ((66, 20, 8), ((64, 4, 3), (68, 16, 7)), 3, 8),            # right x slab (stored q.x[n+2])
    ((66, 20, 8), ((0, 4, 3), (7, 16, 7)), 2, 8),              # left x slab, odd hi
    ((200, 10, 5), None, 16, 8),                               # several row segments along x
]


@pytest.mark.parametrize("n,box,cz,rows", CASES)
def test_fused_thermal3_equals_flux_then_update(oracle, emul3, n, box, cz, rows):
    box = ((0, 0, 0), tuple(x + 2 for x in n)) if box is None else box
    run_case(oracle, emul3, n, box, cz, rows, seed=sum(n) + cz)
