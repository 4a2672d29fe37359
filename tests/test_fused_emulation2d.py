"""
CPU proof of the 2D fused sweeps (chmy.jl_b200/csrc/fused_sv2d.cuh: update_stress! + update_velocity!;
fused_pairs2d.cuh: compute_q! + update_C! and update_thermal_flux! + update_thermal!).

As in test_fused_emulation.py the kernels' phase functions are plain C++ shared by nvcc and the host compiler;
tests/emul/fused_emul2d.cpp runs the 32 lanes of every warp of the launch grid in lock-step.  The result must be
bit-identical to the oracle's first op on the op's whole index range [0, n+1]^2 followed by the second op on the box
(stokes_2d_inc_ve_T.jl:20-60, diffusion_2d.jl:8-19): new values on the box, everything else untouched, the current
buffers of the ping-pong fields read-only.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "fused_emul2d.cpp")
LIB = os.path.join(HERE, "emul", "libfused_emul2d.so")
HDRS = [os.path.join(HERE, "..", "chmy.jl_b200", "csrc", h) for h in ("fused_sv.cuh", "fused_sv2d.cuh", "fused_pairs2d.cuh")]


@pytest.fixture(scope="module")
def emul2():
    from helpers import build_emul
    lib = build_emul("fused_emul2d")
    lib.fused_emul2d_run.restype = C.c_int
    lib.fused_emul2d_pair_run.restype = C.c_int
    return lib


class Pitched2:
    """A 2D field in the library's PITCHED layout (api.cu chmy_field_create): pitch = roundup(sd0, 16), 15-element lead-in."""

    def __init__(self, dense):
        self.sd = dense.shape
        self.pitch = (self.sd[0] + 15) // 16 * 16
        self.lead = 15
        self.flat = np.full(self.lead + self.pitch * self.sd[1] + 32, 777.25)   # slack cells hold junk that must never matter
        self.view()[...] = dense
        self.sy = self.pitch

    def view(self):
        body = self.flat[self.lead:self.lead + self.pitch * self.sd[1]]
        return body.reshape((self.sd[1], self.pitch)).T[:self.sd[0]]

    def p0(self):   # address of logical (0,0) = storage (1,1)
        return self.flat.ctypes.data + 8 * (self.lead + 1 + self.sy)

    def copy(self):
        q = Pitched2.__new__(Pitched2)
        q.__dict__.update(self.__dict__)
        q.flat = self.flat.copy()
        return q


def same(a, b, name):
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    if not ok.all():
        idx = np.argwhere(~ok)
        i = tuple(idx[0])
        raise AssertionError(f"{name}: {len(idx)} cells differ, first at storage {i} (logical {tuple(x - 1 for x in i)}): "
                             f"oracle {a[i]!r} fused {b[i]!r}")


def check_outside(pristine, bufs, sl):
    for k, (orig, buf) in pristine.items():
        a, b = orig.copy(), buf.view().copy()
        a[sl] = 0.0
        b[sl] = 0.0
        same(a, b, "outside-box " + k)


def run_stokes2(o, emul2, n, box, cy, td, fun, seed=0):
    rng = np.random.default_rng(seed)
    g = o.Grid((-1.0, -1.1), (2.0, 2.3), n)
    tau, tau_old, V, rV = o.TensorField(g), o.TensorField(g), o.VectorField(g), o.VectorField(g)
    Pr, dV = o.Field(g, 0), o.Field(g, 0)
    rho = o.Field(g, (0, 1))
    allf = list(tau.values()) + list(tau_old.values()) + list(V.values()) + list(rV.values()) + [Pr, dV, rho]
    for f in allf:
        f.data[...] = rng.random(f.sdims) - 0.5            # interior, halo AND padding
    eta, G, dt = 10.0, 1.3, 0.07
    eta_ve = 1.0 / (1.0 / eta + 1.0 / (G * dt))
    dtau_Pr, dtau_r, nudtau = 0.31, 0.23, 0.011
    inc = o.Inclusion((0, 1), (0.05, -0.1), 0.45, 1.0, 0.25)
    tn, vn = ("xx", "yy", "xy"), ("x", "y")
    named = dict([("t" + c, tau[c]) for c in tn] + [("o" + c, tau_old[c]) for c in tn] + [("V" + c, V[c]) for c in vn] +
                 [("r" + c, rV[c]) for c in vn] + [("Pr", Pr), ("dV", dV), ("rho", rho)])
    cur = {k: Pitched2(f.data) for k, f in named.items()}
    orig = {k: f.data.copy() for k, f in named.items()}
    new = {k: cur[k].copy() for k in ["t" + c for c in tn] + ["V" + c for c in vn] + ["Pr"]}

    o.update_stress(g, (tau, Pr, dV, V, tau_old, eta, eta_ve, G, dt, dtau_Pr, dtau_r), (0, 0), tuple(x + 1 for x in n))
    lo, hi = box
    o.update_velocity(g, (V, rV, Pr, tau, inc if fun else rho, eta_ve, nudtau), lo, tuple(h - 1 for h in hi))

    ptrs = [cur["t" + c].p0() for c in tn] + [cur["o" + c].p0() for c in tn] + [cur["Pr"].p0()] + [cur["V" + c].p0() for c in vn]
    ptrs.append(0 if fun else cur["rho"].p0())
    ptrs += [new["t" + c].p0() for c in tn] + [new["Pr"].p0(), cur["dV"].p0()] + [new["V" + c].p0() for c in vn]
    ptrs += [cur["r" + c].p0() for c in vn]
    assert len(ptrs) == 19
    P = (C.c_void_p * 19)(*ptrs)
    strides = (C.c_int * 4)(cur["Pr"].sy, cur["Vx"].sy, cur["Vy"].sy, cur["txy"].sy)
    assert cur["rho"].sy == cur["Vy"].sy
    bx = (C.c_int * 8)(*lo, *hi, 0, 0, *(x + 2 for x in n))
    sc = (C.c_double * 8)(*g.inv_spacing, eta_ve, dtau_Pr, dtau_r, nudtau, G * dt, eta)
    incv = (C.c_double * 9)(*g.origin, *g.spacing, *inc.c0, inc.r * inc.r, inc.inn, inc.out)
    incloc = (C.c_int * 2)(*inc.loc)
    if int(td) == 2:      # the two-operation division is only ever selected for divisors it is proven exact for
        import chmy_b200
        assert all(chmy_b200.division_two_op_exact(c) for c in (G * dt, eta, eta_ve, 3.0))
    assert emul2.fused_emul2d_run(P, strides, bx, sc, incv, incloc, cy, int(td)) == 0

    sl = tuple(slice(l + 1, h + 1) for l, h in zip(lo, hi))      # logical -> storage index (+1)
    for c in tn:
        same(tau[c].data[sl], new["t" + c].view()[sl], "tau." + c)
    same(Pr.data[sl], new["Pr"].view()[sl], "Pr")
    same(dV.data[sl], cur["dV"].view()[sl], "divV")
    for c in vn:
        same(V[c].data[sl], new["V" + c].view()[sl], "V." + c)
        same(rV[c].data[sl], cur["r" + c].view()[sl], "r_V." + c)
    check_outside({k: (orig[k], new[k] if k in new else cur[k]) for k in orig if k not in ("rho",) and not k.startswith("o")}, None, sl)
    for k in new:                                            # the current buffers are read-only for the kernel
        same(orig[k], cur[k].view(), "current buffer " + k)
    for k in ["o" + c for c in tn] + ["rho"]:
        same(orig[k], cur[k].view(), "read-only " + k)


def run_pair(o, emul2, kind, n, box, cy, seed=0, unroll=1):
    """kind 0: compute_q! + update_C! ; kind 1: update_thermal_flux! + update_thermal!"""
    rng = np.random.default_rng(seed)
    g = o.Grid((-1.0, -1.1), (2.0, 2.3), n)
    Cf, base = o.Field(g, 0), o.Field(g, 0)
    q, V = o.VectorField(g), o.VectorField(g)
    named = {"C": Cf, "base": base, "qx": q["x"], "qy": q["y"], "Vx": V["x"], "Vy": V["y"]}
    for f in named.values():
        f.data[...] = rng.random(f.sdims) - 0.5
    V["x"].data[3, 4] = 0.0
    V["y"].data[5, 2] = -0.0                                  # max(v, 0) / min(v, 0) at signed zeros
    cur = {k: Pitched2(f.data) for k, f in named.items()}
    orig = {k: f.data.copy() for k, f in named.items()}
    new = {"C": cur["C"].copy()}
    coef, dt = 0.7, 0.013
    full_lo, full_hi = (0, 0), tuple(x + 1 for x in n)
    lo, hi = box
    hi_in = tuple(h - 1 for h in hi)
    if kind == 0:
        o.compute_q(g, (q, Cf, coef), full_lo, full_hi)
        o.update_C(g, (Cf, q, dt), lo, hi_in)
    else:
        o.update_thermal_flux(g, (q, Cf, V, coef), full_lo, full_hi)
        o.update_thermal(g, (Cf, base, q, dt), lo, hi_in)
    ptrs = [cur["C"].p0(), new["C"].p0(), cur["base"].p0() if kind else 0, cur["qx"].p0(), cur["qy"].p0(),
            cur["Vx"].p0() if kind else 0, cur["Vy"].p0() if kind else 0]
    P = (C.c_void_p * 7)(*ptrs)
    strides = (C.c_int * 3)(cur["C"].sy, cur["qx"].sy, cur["qy"].sy)
    assert cur["Vx"].sy == cur["qx"].sy and cur["Vy"].sy == cur["qy"].sy and cur["base"].sy == cur["C"].sy
    bx = (C.c_int * 8)(*lo, *hi, 0, 0, *(x + 2 for x in n))
    sc = (C.c_double * 4)(*g.inv_spacing, coef, dt)
    assert emul2.fused_emul2d_pair_run(kind, P, strides, bx, sc, cy, unroll) == 0

    sl = tuple(slice(l + 1, h + 1) for l, h in zip(lo, hi))
    same(Cf.data[sl], new["C"].view()[sl], "C/T")
    same(q["x"].data[sl], cur["qx"].view()[sl], "q.x")
    same(q["y"].data[sl], cur["qy"].view()[sl], "q.y")
    check_outside({"C": (orig["C"], new["C"]), "qx": (orig["qx"], cur["qx"]), "qy": (orig["qy"], cur["qy"])}, None, sl)
    for k in ("C", "base", "Vx", "Vy"):
        same(orig[k], cur[k].view(), "read-only " + k)


CASES2 = [
    # n, box (lo, hi exclusive) or None for the full range, rows per y-chunk
    ((70, 13), None, 4),
    ((70, 13), None, 64),
    ((125, 21), None, 3),
    ((61, 37), None, 5),
    ((9, 5), None, 2),
    ((9, 5), None, 1),
    ((130, 31), ((6, 3), (97, 19)), 7),           # inner region of a split launch
    ((130, 11), ((0, 0), (132, 4)), 8),           # bottom y slab
    ((130, 11), ((0, 9), (132, 13)), 8),          # top y slab
    ((66, 30), ((64, 0), (68, 32)), 16),          # right x slab
    ((66, 30), ((0, 4), (8, 29)), 2),             # left x slab, odd hi
    ((257, 33), None, 128),
]


def _full(n, box):
    return ((0, 0), tuple(x + 2 for x in n)) if box is None else box


@pytest.mark.parametrize("n,box,cy", CASES2)
# td: division mode of the sweep (0 four operations, 1 div.rn.f64, 2 two operations)
@pytest.mark.parametrize("td,fun", [(1, False), (0, True), (0, False), (1, True), (2, True), (2, False)])
def test_fused_stokes2d_equals_stress_then_velocity(oracle, emul2, n, box, cy, td, fun):
    run_stokes2(oracle, emul2, n, _full(n, box), cy, td, fun, seed=sum(n) + cy)


@pytest.mark.parametrize("n,box,cy", CASES2)
@pytest.mark.parametrize("kind", [0, 1], ids=["diffusion", "thermal"])
@pytest.mark.parametrize("unroll", [1, 4])
def test_fused_flux_update_pair_equals_two_ops(oracle, emul2, kind, n, box, cy, unroll):
    """unroll: rows whose operands the kernel requests ahead of the arithmetic (loads of a group are hoisted)."""
    run_pair(oracle, emul2, kind, n, _full(n, box), cy, seed=sum(n) + cy + kind, unroll=unroll)
