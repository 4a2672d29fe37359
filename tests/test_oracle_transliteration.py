"""
CPU-only: pins the oracle's hand-flattened solver arithmetic (oracle/chmy_oracle.c: og_compute_q ... og_update_thermal)
onto a LITERAL transliteration of the reference's kernel bodies.

The reference holds no golden values for its example solvers ("parity unpinned").  What it does hold
is (a) the kernel source text (examples/diffusion_2d.jl:8-19, stokes_2d_inc_ve_T.jl:11-60, stokes_3d_inc_ve_T.jl:11-77)
and (b) known-answer tests for the operators those kernels call, on which the oracle's point functions og_partial /
left / right are pinned (tests/test_oracle_golden.py).  Here every kernel body is re-typed line by line as a Python
expression over those pinned point functions -- same operator calls, same operator precedence and associativity as Julia
parses them (unary minus binds tighter than * and /, `a + b + c + d` folds left, `x -= y*z*w` is x - ((y*z)*w)), Julia's
max/min semantics -- and evaluated point by point over the launch range [0, n+1]^N.  numpy scalars are IEEE numbers with
one rounding per operation in the operands' common type and no contraction, so the two restatements must agree BIT FOR
BIT.  They were written independently (flattened index arithmetic in C vs operator calls in Python); agreement means the
flattening introduced no reassociation, no dropped term and no index slip.

Both element types of the reference's suite (TEST_TYPES = [Float32, Float64], test/common.jl:9): field values, grid numbers
and scalar arguments are numpy scalars of the element type R, the literals of the kernel text (0.5, 2.0, 3.0, 0.0) are
np.float64 -- numpy promotes R op float64 to float64 exactly as Julia promotes Float32 op Float64, and storing into an
R array rounds once, as setindex! on a Float32 array does.  For R = Float64 this is the all-binary64 arithmetic.
"""
import itertools
import math

import numpy as np
import pytest


W = np.float64


def jl_max(a, b):
    """Julia Base.max(x, y) = max(promote(x, y)...): NaN if either is NaN; max(-0.0, +0.0) = +0.0."""
    T = np.result_type(a, b).type
    a, b = T(a), T(b)
    if a != a or b != b:
        return T(math.nan)
    if a == b:
        return b if np.signbit(a) else a
    return a if a > b else b


def jl_min(a, b):
    T = np.result_type(a, b).type
    a, b = T(a), T(b)
    if a != a or b != b:
        return T(math.nan)
    if a == b:
        return a if np.signbit(a) else b
    return a if a < b else b


class K:
    """the operator vocabulary of the kernels, over the oracle's pinned point functions"""

    def __init__(self, o, g):
        self.o, self.g, self.R = o, g, g.dtype.type

    def d(self, dim, f, I):                      # ∂x / ∂y / ∂z (cartesian_field_operators.jl:17-46 -> field_operators.jl:20-24)
        return self.R(self.o.partial(self.g, f, dim, *I))

    def left(self, dim, f, I):                   # leftx ... (GridOperators.jl:23-33, field_operators.jl:2-6)
        J = list(I)
        if f.loc[dim] == 0:
            J[dim] -= 1
        return self.R(f.at(*J))

    def right(self, dim, f, I):                  # rightx ... (field_operators.jl:8-12)
        J = list(I)
        if f.loc[dim] == 1:
            J[dim] += 1
        return self.R(f.at(*J))

    def divg(self, V, I):                        # field_operators.jl:50-55: @ncall N (+) -> left fold
        s = self.d(0, V[0], I)
        for D in range(1, len(V)):
            s = s + self.d(D, V[D], I)
        return s


def rnd(o, g, loc, rng):
    f = o.Field(g, loc)
    f.data[...] = rng.random(f.sdims) - 0.5                    # interior, halo AND padding
    return f


def launch_range(n):
    return itertools.product(*[range(0, x + 2) for x in n])    # Launcher: I = J - 1 over worksize n+2 (KernelLaunch.jl:41,109)


def put(f, I, v):
    f.data[tuple(i + 1 for i in I)] = v


def same(a, b, what):
    ok = (a.data == b.data) | (np.isnan(a.data) & np.isnan(b.data))
    assert ok.all(), f"{what}: {int((~ok).sum())} cells differ, first at storage {tuple(np.argwhere(~ok)[0])}"


def clone(o, f):
    c = o.Field(f.grid, f.loc)
    c.data[...] = f.data
    return c


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_diffusion_kernels_literally(oracle, dtype):
    o = oracle
    n = (7, 6)
    g = o.Grid((-1.0, -1.0), (2.0, 2.3), n, dtype=dtype)
    R = g.dtype.type
    rng = np.random.default_rng(1)
    k = K(o, g)
    C = rnd(o, g, 0, rng)
    q = {c: rnd(o, g, l, rng) for c, l in (("x", (1, 0)), ("y", (0, 1)))}
    C2, q2 = clone(o, C), {c: clone(o, f) for c, f in q.items()}
    chi, dt = R(0.83), R(0.0137)
    L = o.Launcher(g)
    o.launch(L, g, o.compute_q, (q, C, chi))
    for I in launch_range(n):                                  # diffusion_2d.jl:8-13
        put(q2["x"], I, -chi * k.d(0, C2, I))
        put(q2["y"], I, -chi * k.d(1, C2, I))
    same(q["x"], q2["x"], "compute_q! q.x"); same(q["y"], q2["y"], "compute_q! q.y")
    o.launch(L, g, o.update_C, (C, q, dt))
    for I in launch_range(n):                                  # :15-19   C[I] -= Δt * divg(q, g, I)
        put(C2, I, R(C2.at(*I)) - dt * k.divg([q2["x"], q2["y"]], I))
    same(C, C2, "update_C!")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [(7, 6), (6, 5, 4)])
def test_stokes_kernels_literally(oracle, n, dtype):
    o, nd = oracle, len(n)
    g = o.Grid((-1.0,) * nd, tuple(2.0 + 0.3 * d for d in range(nd)), n, dtype=dtype)
    R = g.dtype.type
    rng = np.random.default_rng(2)
    k = K(o, g)
    vn = "xyz"[:nd]
    tn = ("xx", "yy", "xy") if nd == 2 else ("xx", "yy", "zz", "xy", "xz", "yz")
    tau, tau_old, V, rV, qT = o.TensorField(g), o.TensorField(g), o.VectorField(g), o.VectorField(g), o.VectorField(g)
    Pr, dV, T, To = (o.Field(g, 0) for _ in range(4))
    rho = o.Field(g, tuple(1 if d == nd - 1 else 0 for d in range(nd)))
    every = list(tau.values()) + list(tau_old.values()) + list(V.values()) + list(rV.values()) + list(qT.values()) + [Pr, dV, T, To, rho]
    for f in every:
        f.data[...] = rng.random(f.sdims) - 0.5
    V["x"].data[2, 3] = 0.0
    V["y"].data[3, 2] = -0.0                                   # max(v, 0.0) / min(v, 0.0) at signed zeros
    V[vn[-1]].data[4, 4] = math.nan
    c = lambda d: {kk: clone(o, f) for kk, f in d.items()}
    tau2, tauo2, V2, rV2, qT2 = c(tau), c(tau_old), c(V), c(rV), c(qT)
    Pr2, dV2, T2, To2 = clone(o, Pr), clone(o, dV), clone(o, T), clone(o, To)
    eta, G, dt, dPr, dr, nud, lam = (R(x) for x in (10.0, 1.3, 0.0171, 0.0213, 0.613, 0.00931, 3.3e-4))
    eta_ve = R(1.0 / (1.0 / float(eta) + 1.0 / (float(G) * float(dt))))        # a host scalar of the driver, passed as R
    L = o.Launcher(g)
    Vl = lambda d: [d[cc] for cc in vn]

    # ---- update_old! (stokes_3d_inc_ve_T.jl:11-21)
    o.launch(L, g, o.update_old, (T, tau, To, tau_old))
    for I in launch_range(n):
        put(To2, I, R(T2.at(*I)))
        for cc in tn:
            put(tauo2[cc], I, R(tau2[cc].at(*I)))
    same(To, To2, "update_old! T_old")
    for cc in tn:
        same(tau_old[cc], tauo2[cc], "update_old! tau_old." + cc)
    for f, f2 in list(zip(tau_old.values(), tauo2.values())):  # independent old stresses for the stress kernel
        f.data[...] = rng.random(f.sdims) - 0.5
        f2.data[...] = f.data

    # ---- update_stress! (:23-46; 2D stokes_2d_inc_ve_T.jl:20-34)
    o.launch(L, g, o.update_stress, (tau, Pr, dV, V, tau_old, eta, eta_ve, G, dt, dPr, dr))
    dims = {"x": 0, "y": 1, "z": 2}
    for I in launch_range(n):
        e = {}
        for a in vn:
            e[a + a] = k.d(dims[a], V2[a], I)                                            # ε̇xx = ∂x(V.x, g, I...)
        for ab in tn[nd:]:
            a, b = ab
            e[ab] = W(0.5) * (k.d(dims[b], V2[a], I) + k.d(dims[a], V2[b], I))           # ε̇xy = 0.5 * (∂y(V.x) + ∂x(V.y))
        put(dV2, I, k.divg(Vl(V2), I))                                                   # ∇V[I...] = divg(V, g, I...)
        div = R(dV2.at(*I))
        put(Pr2, I, R(Pr2.at(*I)) - div * eta_ve * dPr)                                  # Pr[I...] -= ∇V[I...] * η_ve * dτ_Pr
        r = {}
        for cc in tn:
            t, to = R(tau2[cc].at(*I)), R(tauo2[cc].at(*I))
            if cc[0] == cc[1]:   # r_τxx = -(τ.xx - τ_old.xx) / (G * dt) - τ.xx / η + 2.0 * (ε̇xx - ∇V / 3.0)
                r[cc] = -(t - to) / (G * dt) - t / eta + W(2.0) * (e[cc] - div / W(3.0))
            else:                # r_τxy = -(τ.xy - τ_old.xy) / (G * dt) - τ.xy / η + 2.0 * ε̇xy
                r[cc] = -(t - to) / (G * dt) - t / eta + W(2.0) * e[cc]
            assert type(r[cc]) is W                                                      # the literals promoted the residual
        for cc in tn:
            put(tau2[cc], I, R(tau2[cc].at(*I)) + r[cc] * eta_ve * dr)                   # τ.xx[I...] += r_τxx * η_ve * dτ_r
    same(dV, dV2, "update_stress! divV"); same(Pr, Pr2, "update_stress! Pr")
    for cc in tn:
        same(tau[cc], tau2[cc], "update_stress! tau." + cc)

    # ---- update_velocity! (:48-57; 2D :36-43)
    o.launch(L, g, o.update_velocity, (V, rV, Pr, tau, rho, eta_ve, nud))
    for I in launch_range(n):
        if nd == 3:
            rx = -k.d(0, Pr2, I) + k.d(0, tau2["xx"], I) + k.d(1, tau2["xy"], I) + k.d(2, tau2["xz"], I)
            ry = -k.d(1, Pr2, I) + k.d(1, tau2["yy"], I) + k.d(0, tau2["xy"], I) + k.d(2, tau2["yz"], I)
            rz = -k.d(2, Pr2, I) + k.d(2, tau2["zz"], I) + k.d(0, tau2["xz"], I) + k.d(1, tau2["yz"], I) - R(rho.at(*I))
            res = {"x": rx, "y": ry, "z": rz}
        else:
            rx = -k.d(0, Pr2, I) + k.d(0, tau2["xx"], I) + k.d(1, tau2["xy"], I)
            ry = -k.d(1, Pr2, I) + k.d(1, tau2["yy"], I) + k.d(0, tau2["xy"], I) - R(rho.at(*I))
            res = {"x": rx, "y": ry}
        for a in vn:
            put(rV2[a], I, res[a])
        for a in vn:
            put(V2[a], I, R(V2[a].at(*I)) + R(rV2[a].at(*I)) * nud / eta_ve)              # V.x[I...] += r_V.x[I...] * νdτ / η_ve
    for a in vn:
        same(rV[a], rV2[a], "update_velocity! r_V." + a)
        same(V[a], V2[a], "update_velocity! V." + a)

    # ---- update_thermal_flux! (:59-71) and update_thermal! (:73-77)
    o.launch(L, g, o.update_thermal_flux, (qT, T, V, lam))
    for I in launch_range(n):
        for a in vn:
            D = dims[a]
            v = R(V2[a].at(*I))
            put(qT2[a], I, -lam * k.d(D, T2, I) + jl_max(v, W(0.0)) * k.left(D, T2, I) + jl_min(v, W(0.0)) * k.right(D, T2, I))
    for a in vn:
        same(qT[a], qT2[a], "update_thermal_flux! qT." + a)
    o.launch(L, g, o.update_thermal, (T, To, qT, dt))
    for I in launch_range(n):
        put(T2, I, R(To2.at(*I)) - dt * k.divg(Vl(qT2), I))
    same(T, T2, "update_thermal!")
