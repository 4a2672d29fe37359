"""
CPU proof of the grid-operator kernels (chmy.jl_b200/csrc/operators.cuh, CHMY_OP_OPERATOR).

The operator point functions are plain C++ shared by nvcc and the host compiler; tests/emul/operators_emul.cpp applies
them over the launch range [0, n+1]^N on host arrays in the library's PITCHED layout.  Every operator, for every
staggered location of its operands and in 1D / 2D / 3D, must be bit-identical to the oracle's independent restatement
(oracle/chmy_oracle.c: og_apply_operator, itself pinned on test/test_grid_operators.jl and test/test_interpolations.jl
by tests/test_oracle_golden.py).  The -m gpu suite then checks the compiled kernel against the same oracle.
"""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "operators_emul.cpp")
LIB = os.path.join(HERE, "emul", "liboperators_emul.so")
HDR = os.path.join(HERE, "..", "chmy.jl_b200", "csrc", "operators.cuh")


class OprField(C.Structure):
    _fields_ = [("p", C.c_void_p), ("sy", C.c_longlong), ("sz", C.c_longlong), ("loc", C.c_int * 3)]


class OprArgs(C.Structure):
    _fields_ = [("oper", C.c_int), ("dim", C.c_int), ("nd", C.c_int), ("ndst", C.c_int), ("dst", OprField * 3),
                ("a", OprField * 3), ("k", OprField), ("id", C.c_double * 3)]


class OprArgsF32(C.Structure):
    _fields_ = [("oper", C.c_int), ("dim", C.c_int), ("nd", C.c_int), ("ndst", C.c_int), ("dst", OprField * 3),
                ("a", OprField * 3), ("k", OprField), ("id", C.c_float * 3)]


DTYPES = [np.float64, np.float32]            # TEST_TYPES of the reference (test/common.jl:9)


@pytest.fixture(scope="module")
def emul():
    from helpers import build_emul
    lib = build_emul("operators_emul")
    assert lib.operators_emul_sizeof_args(0) == C.sizeof(OprArgs) and lib.operators_emul_sizeof_args(1) == C.sizeof(OprArgsF32)
    return lib


class Pitched:
    """Host copy of an oracle field in the library's PITCHED layout (api.cu chmy_field_create)."""

    def __init__(self, of):
        sd = tuple(of.sdims) + (1,) * (3 - of.nd)
        self.nd, self.sd, self.loc = of.nd, sd, tuple(of.loc) + (0,) * (3 - of.nd)
        self.es = of.data.dtype.itemsize
        per128 = 128 // self.es                                  # 16 doubles | 32 floats per 128-byte line
        self.pitch = (sd[0] + per128 - 1) // per128 * per128
        self.lead = per128 - 1
        self.flat = np.full(self.lead + self.pitch * sd[1] * sd[2] + 2 * per128, 777.25, dtype=of.data.dtype)  # slack = junk
        self.view()[...] = of.data.reshape(sd, order="F")
        self.sy = self.pitch if of.nd > 1 else 0
        self.sz = self.pitch * sd[1] if of.nd > 2 else 0

    def view(self):
        sd = self.sd
        body = self.flat[self.lead:self.lead + self.pitch * sd[1] * sd[2]]
        return body.reshape((sd[2], sd[1], self.pitch)).transpose(2, 1, 0)[:sd[0]]

    def opr(self):
        f = OprField()
        # logical 0 of every active dim = storage index 1 (field.jl:18 with halo 1)
        f.p = self.flat.ctypes.data + self.es * (self.lead + 1 + self.sy + self.sz)
        f.sy, f.sz = self.sy, self.sz
        for a in range(3):
            f.loc[a] = self.loc[a]
        return f

    def dense(self, nd):
        return np.asfortranarray(self.view()).reshape(self.sd[:nd], order="F")


def flip(loc, d):
    return tuple(1 - l if a == d else l for a, l in enumerate(loc))


def run_both(o, emul, g, kind, dst_o, src_o, k_o=None, dim=0):
    """Apply `kind` through the oracle (in place on dst_o) and through the emulation (on pitched copies); compare bits."""
    dst_o = dst_o if isinstance(dst_o, list) else [dst_o]
    src_o = src_o if isinstance(src_o, list) else [src_o]
    pd = [Pitched(f) for f in dst_o]
    ps = [Pitched(f) for f in src_o]
    pk = Pitched(k_o) if k_o is not None else None
    f32 = g.dtype == np.float32
    a = OprArgsF32() if f32 else OprArgs()
    a.oper, a.dim, a.nd, a.ndst = o.OPER[kind], dim, g.nd, len(pd)
    for c, f in enumerate(pd):
        a.dst[c] = f.opr()
    for c, f in enumerate(ps):
        a.a[c] = f.opr()
    if pk is not None:
        a.k = pk.opr()
    for d in range(g.nd):
        a.id[d] = g.inv_spacing[d]
    lo = (C.c_int * 3)(0, 0, 0)
    hi = (C.c_int * 3)(*([n + 1 for n in g.n] + [0] * (3 - g.nd)))
    assert (emul.operators_emul_run_f32 if f32 else emul.operators_emul_run)(C.byref(a), lo, hi) == 0
    o.apply_operator(g, kind, dst_o, src_o, k=k_o, dim=dim)
    for c, (fo, fp) in enumerate(zip(dst_o, pd)):
        got = fp.dense(g.nd)
        same = (fo.data == got) | (np.isnan(fo.data) & np.isnan(got))
        assert same.all(), f"{kind} dim={dim} dst{c} loc={fo.loc}: {np.argwhere(~same)[:3]}"
        assert np.abs(fo.data).max() > 0
    for f, fp in zip(src_o, ps):                       # sources untouched
        assert np.array_equal(f.data, fp.dense(g.nd))


def rnd_field(o, g, loc, rng, positive=False):
    f = o.Field(g, loc)
    f.data[...] = rng.random(f.sdims) + 0.5 if positive else rng.random(f.sdims) - 0.5      # interior, halo AND padding
    assert f.data.dtype == g.dtype
    return f


GRIDS = [((-1.0,), (2.0,), (9,)), ((-1.0, 0.5), (2.0, 1.7), (7, 5)), ((-5.0, -5.0, -5.0), (10.0, 9.0, 8.0), (6, 5, 4))]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS)
def test_single_field_operators_every_location(oracle, emul, origin, extent, n, dtype):
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(3)
    for loc in itertools.product((0, 1), repeat=nd):
        f = rnd_field(o, g, loc, rng)
        for dim in range(nd):
            for kind in ("left", "right", "delta", "partial"):
                run_both(o, emul, g, kind, rnd_field(o, g, flip(loc, dim), rng), f, dim=dim)
            run_both(o, emul, g, "partial2", rnd_field(o, g, loc, rng), f, dim=dim)
            for kloc in itertools.product((0, 1), repeat=nd):
                run_both(o, emul, g, "dkd", rnd_field(o, g, loc, rng), f, k_o=rnd_field(o, g, kloc, rng), dim=dim)
        run_both(o, emul, g, "lapl", rnd_field(o, g, loc, rng), f)
        for kloc in itertools.product((0, 1), repeat=nd):
            run_both(o, emul, g, "divg_grad", rnd_field(o, g, loc, rng), f, k_o=rnd_field(o, g, kloc, rng))
        fpos = rnd_field(o, g, loc, rng, positive=True)
        for to in itertools.product((0, 1), repeat=nd):
            run_both(o, emul, g, "lerp", rnd_field(o, g, to, rng), f)
            run_both(o, emul, g, "hlerp", rnd_field(o, g, to, rng), fpos)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS)
def test_vector_operators(oracle, emul, origin, extent, n, dtype):
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(4)
    ctr = (0,) * nd
    V = [rnd_field(o, g, flip(ctr, d), rng) for d in range(nd)]               # VectorField locations (field.jl:148)
    run_both(o, emul, g, "divg", rnd_field(o, g, ctr, rng), V)
    run_both(o, emul, g, "vmag", rnd_field(o, g, ctr, rng), V)
    # divg of a "dual" vector field (components at Center along their own axis) lives at Vertex
    vtx = (1,) * nd
    W = [rnd_field(o, g, flip(vtx, d), rng) for d in range(nd)]
    run_both(o, emul, g, "divg", rnd_field(o, g, vtx, rng), W)
    for floc in (ctr, vtx):
        f = rnd_field(o, g, floc, rng)
        run_both(o, emul, g, "grad", [rnd_field(o, g, flip(floc, d), rng) for d in range(nd)], f)
        for kloc in itertools.product((0, 1), repeat=nd):
            run_both(o, emul, g, "kgrad", [rnd_field(o, g, flip(floc, d), rng) for d in range(nd)], f,
                     k_o=rnd_field(o, g, kloc, rng))


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_identities_through_the_kernel_functions(oracle, emul, dtype):
    """test/test_grid_operators.jl:21-61 with the operator kernels' own code: divg(grad C) == sum of the three partial
    derivatives with `==` (:41) and lapl == sum of second derivatives with `==` (:61)."""
    o = oracle
    g = o.Grid((-5.0, -5.0, -5.0), (10.0, 10.0, 10.0), (12, 10, 8), dtype=dtype)
    Ci = o.Field(g, 0)
    Ci.set_fun(lambda x, y, z: np.exp(-x ** 2 - y ** 2 - z ** 2))
    V = [o.Field(g, flip((0, 0, 0), d)) for d in range(3)]
    run_both(o, emul, g, "grad", V, Ci)                       # V now holds the oracle's result (== the emulation's)
    C2 = o.Field(g, 0)
    run_both(o, emul, g, "divg", C2, V)
    parts = []
    for d in range(3):
        P = o.Field(g, 0)
        run_both(o, emul, g, "partial", P, V[d], dim=d)
        parts.append(P.interior().copy())
    assert np.array_equal(C2.interior(), (parts[0] + parts[1]) + parts[2])
    L = o.Field(g, 0)
    run_both(o, emul, g, "lapl", L, Ci)
    parts = []
    for d in range(3):
        P = o.Field(g, 0)
        run_both(o, emul, g, "partial2", P, Ci, dim=d)
        parts.append(P.interior().copy())
    assert np.array_equal(L.interior(), (parts[0] + parts[1]) + parts[2])
    assert np.abs(L.interior()).max() > 1e-3
