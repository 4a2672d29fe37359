"""
Parity of the CUDA path (through the C ABI, via the host mirror) against the CPU oracle.  GPU only.
Tolerances: integer/index/byte work (pack/unpack, BC index sets) bit-exact; Float64 fields: the north_star allows
1e-12 relative -- the kernels are written to be bit-identical to the oracle, and these tests assert exactly that
(tol = 0) wherever no division by a run-time scalar takes a different code path.
"""
import math

import numpy as np
import pytest

from helpers import assert_same, fill_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    return chmy_b200


@pytest.fixture(scope="module")
def arch(ch):
    a = ch.Arch(ch.B200Backend())
    yield a
    a.close()


LOCS = {0: "Center", 1: "Vertex"}


def mk_grids(ch, o, arch, n, origin=None, extent=None, dtype=np.float64):
    nd = len(n)
    origin = origin or tuple(-1.0 - 0.1 * d for d in range(nd))
    extent = extent or tuple(2.0 + 0.3 * d for d in range(nd))
    return o.Grid(origin, extent, n, dtype=dtype), ch.UniformGrid(arch, origin=origin, extent=extent, dims=n, dtype=dtype)


DTYPES = [np.float64, np.float32]        # TEST_TYPES of the reference's suite (test/common.jl:9)


def bloc(ch, loc):
    return tuple(ch.Vertex() if l else ch.Center() for l in loc)


# ------------------------------------------------------------------------------------------------ fields
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,loc", [((7,), (1,)), ((33, 18), (0, 1)), ((12, 10, 8), (1, 0, 1)), ((5, 3, 2), (0, 0, 0))])
def test_field_roundtrip_and_layout(ch, arch, oracle, n, loc, layout):
    og, bg = mk_grids(ch, oracle, arch, n)
    of = oracle.Field(og, loc)
    bf = ch.Field(arch, bg, bloc(ch, loc), layout=layout)
    assert bf.dims == of.dims
    assert np.array_equal(bf.parent(), np.zeros(of.sdims))              # zero-initialised (field.jl:59)
    rng = np.random.default_rng(1)
    fill_pair(rng, of, bf)
    assert_same(of, bf, "roundtrip")
    assert np.array_equal(ch.interior(bf), of.interior())
    assert np.array_equal(ch.interior(bf, with_halo=True), of.interior(with_halo=True))
    info = bf.info()
    assert info.stride[0] == 1 and info.layout == layout
    if layout == 1 and len(n) > 1:
        assert info.stride[1] == of.sdims[0]
    if layout == 0:
        assert (info.origin_ptr - 8) % 128 == 0                         # logical index 0 of each row is 128B aligned
    # set!(f, val) touches the interior only
    ch.set_(bf, 3.5)
    of.set(3.5)
    assert_same(of, bf, "set scalar")
    A = rng.random(of.dims)
    ch.set_(bf, A)
    of.set(A)
    assert_same(of, bf, "set array")
    assert ch.maxabs(bf) == of.maxabs()
    ch.fill_parent_(bf, float("nan"))
    assert np.isnan(bf.parent()).all()


def test_set_continuous_known_answers(ch, arch):
    """test/test_fields.jl:19-55 through the CUDA path."""
    g = ch.UniformGrid(arch, origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0), dims=(2, 2, 2))
    f = ch.Field(arch, g, (ch.Center(), ch.Vertex(), ch.Center()))
    ch.fill_parent_(f, float("nan"))
    ch.set_(f, g, lambda x, y, z: y)
    exp_y = np.zeros((2, 3, 2)); exp_y[:, 1, :], exp_y[:, 2, :] = 0.5, 1.0
    assert np.array_equal(ch.interior(f), exp_y)
    ch.set_(f, g, lambda x, y, z: x)
    exp_x = np.zeros((2, 3, 2)); exp_x[0], exp_x[1] = 0.25, 0.75
    assert np.array_equal(ch.interior(f), exp_x)
    ch.set_(f, g, lambda x, y, z, sc: y * sc, parameters=(2.0,))
    assert np.array_equal(ch.interior(f), 2 * exp_y)
    assert np.isnan(f.parent()[0]).all()                                # halo/padding untouched


@pytest.mark.parametrize("n", [(40, 27), (19, 14, 11)])
def test_set_inclusion_device_vs_oracle(ch, arch, oracle, n):
    og, bg = mk_grids(ch, oracle, arch, n, origin=(-1.0,) * len(n), extent=(2.0,) * len(n))
    nd = len(n)
    for loc in [(0,) * nd, tuple(1 if d == nd - 1 else 0 for d in range(nd))]:
        of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
        oracle.set_inclusion(of, oracle.Inclusion(loc, (0.05,) * nd, 0.31, 1.0, 0.1))
        par = dict(zip(("x0", "y0", "z0"), (0.05,) * nd))
        ch.set_(bf, bg, ch.init_incl, parameters={**par, "r": 0.31, "in": 1.0, "out": 0.1})
        assert_same(of, bf, "inclusion")
        assert 0 < (of.interior() == 1.0).sum() < of.interior().size


def test_maxabs_nan_and_exactness(ch, arch, oracle):
    og, bg = mk_grids(ch, oracle, arch, (65, 33, 9))
    of, bf = oracle.Field(og, 0), ch.Field(arch, bg, ch.Center())
    rng = np.random.default_rng(5)
    a = fill_pair(rng, of, bf, scale=1e-3)
    assert ch.maxabs(bf) == np.abs(a[2:-2, 2:-2, 2:-2]).max() == of.maxabs()
    assert ch.maxabs(bf, with_halo=True) == np.abs(a[1:-1, 1:-1, 1:-1]).max()
    a[10, 10, 5] = np.nan
    bf.from_host(a, [-1] * 3, [d + 2 for d in of.dims])
    assert math.isnan(ch.maxabs(bf))                                    # Julia's maximum propagates NaN


# ------------------------------------------------------------------------------------------------ BCs
BC_CASES = [((8,), (0,)), ((8,), (1,)), ((8, 8), (0, 1)), ((8, 8, 6), (0, 1, 0)), ((13, 7, 5), (1, 1, 0))]


@pytest.mark.parametrize("n,loc", BC_CASES)
def test_bc_matches_oracle_bit_exact(ch, arch, oracle, n, loc):
    """Dirichlet/Neumann x homogeneous/valued on random full arrays: every touched cell (incl. edges, corners and
    padding rows, where the z->y->x order matters) must match the oracle exactly."""
    nd = len(n)
    og, bg = mk_grids(ch, oracle, arch, n, origin=(-math.pi,) * nd, extent=(2 * math.pi,) * nd)
    of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    rng = np.random.default_rng(7)
    for mk_o, mk_b in [(oracle.Dirichlet, ch.Dirichlet), (oracle.Neumann, ch.Neumann)]:
        for val in (None, 2.0):
            fill_pair(rng, of, bf)
            oracle.bc_(og, (of, mk_o(val)))
            ch.bc_(arch, bg, (bf, mk_b(val)))
            assert_same(of, bf, f"bc {mk_o.__name__}({val})")
    # mixed per-axis spec with different left/right conditions
    if nd >= 2:
        fill_pair(rng, of, bf)
        so = {"x": (oracle.Dirichlet(1.5), oracle.Neumann(-0.5)), "y": oracle.Neumann()}
        sb = {"x": (ch.Dirichlet(1.5), ch.Neumann(-0.5)), "y": ch.Neumann()}
        oracle.bc_(og, (of, so))
        ch.bc_(arch, bg, (bf, sb))
        assert_same(of, bf, "bc mixed")


@pytest.mark.parametrize("n,loc", [((8, 8), (0, 1)), ((9, 6), (1, 0)), ((8, 8, 6), (0, 1, 0)), ((13, 7, 5), (1, 1, 0))])
def test_valued_bc_field_and_boundary_function_bit_exact(ch, arch, oracle, n, loc):
    """Field-valued (first_order_boundary_condition.jl:38-40) and BoundaryFunction-valued (boundary_function.jl:28-44)
    Dirichlet / Neumann conditions, every dim and side, against the oracle on random full arrays."""
    o = oracle
    nd = len(n)
    og, bg = mk_grids(ch, o, arch, n, origin=(-math.pi,) * nd, extent=(2 * math.pi,) * nd)
    of, bf = o.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    rng = np.random.default_rng(17)
    names = ("x", "y", "z")[:nd]
    for D, ax in enumerate(names):
        # lower-dimensional value fields on the transverse grid
        tg_o = o.transverse_grid(og, D)
        keep = [a for a in range(nd) if a != D]
        tg_b = ch.UniformGrid(arch, origin=[og.origin[a] for a in keep], extent=[og.extent[a] for a in keep],
                              dims=[n[a] for a in keep])
        vo, vb = o.Field(tg_o, o.VERTEX), ch.Field(arch, tg_b, ch.Vertex())
        fill_pair(rng, vo, vb)
        for mk_o, mk_b in [(o.Dirichlet, ch.Dirichlet), (o.Neumann, ch.Neumann)]:
            fill_pair(rng, of, bf)
            o.bc_(og, (of, {ax: mk_o(vo)}))
            ch.bc_(arch, bg, (bf, {ax: mk_b(vb)}))
            assert_same(of, bf, f"field-valued {mk_o.__name__} along {ax}")
            fun = (lambda *x: math.cos(x[0]) * (1.0 + (x[1] if len(x) > 2 else 0.0)) + x[-1])
            fo = o.BoundaryFunction(fun, parameters=(0.25,))
            fb = ch.BoundaryFunction(fun, parameters=(0.25,))
            fill_pair(rng, of, bf)
            o.bc_(og, (of, {ax: (mk_o(fo), mk_o(2.0))}))
            ch.bc_(arch, bg, (bf, {ax: (mk_b(fb), mk_b(2.0))}))
            assert_same(of, bf, f"function-valued {mk_o.__name__} along {ax}")
            dfo = o.BoundaryFunction(lambda g, l, dim, *I: float(sum(I)) * 0.5, discrete=True, reduce_dims=False)
            dfb = ch.BoundaryFunction(lambda g, l, dim, *I: float(sum(I)) * 0.5, discrete=True, reduce_dims=False)
            fill_pair(rng, of, bf)
            o.bc_(og, (of, {ax: mk_o(dfo)}))
            ch.bc_(arch, bg, (bf, {ax: mk_b(dfb)}))
            assert_same(of, bf, f"discrete-function-valued {mk_o.__name__} along {ax}")


def test_bc_known_answers_cuda(ch, arch):
    """test/test_boundary_conditions.jl:143-205 (3D, (C,V,C)) straight on the CUDA path."""
    n = (8, 8, 6)
    g = ch.UniformGrid(arch, origin=(-math.pi,) * 3, extent=(2 * math.pi,) * 3, dims=n)
    f = ch.Field(arch, g, (ch.Center(), ch.Vertex(), ch.Center()))
    loc = (0, 1, 0)

    def face(a, d, idx):
        sl = [slice(1, -1)] * 3
        sl[d] = idx
        return a[tuple(sl)]

    ch.set_(f, 1.0); ch.bc_(arch, g, (f, ch.Dirichlet()))
    a = ch.interior(f, with_halo=True)
    for d in range(3):
        if loc[d] == 0:
            assert np.allclose(face(a, d, 0), -face(a, d, 1)) and np.allclose(face(a, d, -1), -face(a, d, -2))
        else:
            assert np.allclose(face(a, d, 1), 0.0) and np.allclose(face(a, d, -2), 0.0)
    ch.set_(f, 1.0); ch.bc_(arch, g, (f, ch.Neumann(2.0)))
    a = ch.interior(f, with_halo=True)
    for d in range(3):
        h = ch.spacing(g)[d]
        assert np.allclose((face(a, d, 1) - face(a, d, 0)) / h, 2.0)
        assert np.allclose((face(a, d, -1) - face(a, d, -2)) / h, 2.0)


# ------------------------------------------------------------------------------------------------ halo slabs
@pytest.mark.parametrize("n,loc", [((9, 6), (1, 0)), ((9, 6), (0, 0)), ((11, 7, 5), (1, 0, 0)), ((11, 7, 5), (0, 1, 1))])
def test_halo_pack_unpack_bit_exact(ch, arch, oracle, n, loc):
    """communication_views.jl:1-34: index-encoded fields (catches orientation errors), every (dim, side)."""
    import ctypes as C
    from chmy_b200 import _lib as L
    og, bg = mk_grids(ch, oracle, arch, n)
    of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    enc = np.arange(of.data.size, dtype=np.float64).reshape(of.sdims, order="F") + 1e6
    of.data[...] = enc
    bf.from_host(enc, [-1] * len(n), [d + 2 for d in of.dims])
    for D in range(len(n)):
        for S in range(2):
            ref = oracle.pack_send(of, D, S)
            ln = C.c_int64()
            L.check(L.lib().chmy_halo_slab_len(bf.handle, D, C.byref(ln)))
            assert ln.value == ref.size
            buf = np.empty(ref.size)
            L.check(L.lib().chmy_halo_pack(arch.ctx, bf.handle, D, S, buf.ctypes.data_as(C.c_void_p)))
            assert np.array_equal(buf, ref), (D, S)
            # unpack a recognisable slab into the recv position
            msg = -ref[::-1].copy()
            oracle.unpack_recv(of, D, S, msg)
            L.check(L.lib().chmy_halo_unpack(arch.ctx, bf.handle, D, S, msg.ctypes.data_as(C.c_void_p)))
            assert_same(of, bf, f"unpack {D},{S}")


# ------------------------------------------------------------------------------------------------ single ops
def _stokes_pair(ch, o, arch, n, rng, dtype=np.float64):
    og, bg = mk_grids(ch, o, arch, n, origin=(-1.0,) * len(n), extent=(2.0,) * len(n), dtype=dtype)
    O = dict(tau=o.TensorField(og), tau_old=o.TensorField(og), V=o.VectorField(og), rV=o.VectorField(og),
             qT=o.VectorField(og), Pr=o.Field(og, 0), dV=o.Field(og, 0), T=o.Field(og, 0), To=o.Field(og, 0))
    B = dict(tau=ch.TensorField(arch, bg), tau_old=ch.TensorField(arch, bg), V=ch.VectorField(arch, bg),
             rV=ch.VectorField(arch, bg), qT=ch.VectorField(arch, bg), Pr=ch.Field(arch, bg), dV=ch.Field(arch, bg),
             T=ch.Field(arch, bg), To=ch.Field(arch, bg))
    rl = tuple(1 if d == len(n) - 1 else 0 for d in range(len(n)))
    O["rho"], B["rho"] = o.Field(og, rl), ch.Field(arch, bg, bloc(ch, rl))
    pairs = []
    for k in O:
        if isinstance(O[k], dict):
            for c in O[k]:
                pairs.append((f"{k}.{c}", O[k][c], getattr(B[k], c)))
        else:
            pairs.append((k, O[k], B[k]))
    for _, a, b in pairs:
        fill_pair(rng, a, b)
    return og, bg, O, B, pairs


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [(30, 22, 14), (17, 9, 5), (65, 33, 4), (126, 126), (37, 258)])
def test_every_op_bit_exact(ch, arch, oracle, n, dtype):
    """every solver kernel against the oracle, bit for bit, in both element types: with Float32 fields the Float64 literals
    of the kernels promote exactly the sub-expressions Julia would promote (ops.cu `W`, chmy_oracle.h og_wide)"""
    o = oracle
    rng = np.random.default_rng(11)
    og, bg, O, B, pairs = _stokes_pair(ch, o, arch, n, rng, dtype)
    Lo, Lb = o.Launcher(og), ch.Launcher(arch, bg)
    sc = {k: dtype(v) for k, v in dict(eta=10.0, eta_ve=0.737, G=1.3, dt=0.0171, dPr=0.0213, dr=0.613, nud=0.00931, lam=3.3e-4).items()}

    def check(tag):
        for name, a, b in pairs:
            assert_same(a, b, f"{tag}:{name}")

    o.launch(Lo, og, o.update_old, (O["T"], O["tau"], O["To"], O["tau_old"]))
    Lb(arch, bg, (ch.update_old_, (B["T"], B["tau"], B["To"], B["tau_old"])))
    check("update_old")
    for _, a, b in pairs:
        fill_pair(rng, a, b)
    o.launch(Lo, og, o.update_stress, (O["tau"], O["Pr"], O["dV"], O["V"], O["tau_old"], sc["eta"], sc["eta_ve"], sc["G"],
                                       sc["dt"], sc["dPr"], sc["dr"]))
    Lb(arch, bg, (ch.update_stress_, (B["tau"], B["Pr"], B["dV"], B["V"], B["tau_old"], sc["eta"], sc["eta_ve"], sc["G"],
                                      sc["dt"], sc["dPr"], sc["dr"], bg)))
    check("update_stress")
    o.launch(Lo, og, o.update_velocity, (O["V"], O["rV"], O["Pr"], O["tau"], O["rho"], sc["eta_ve"], sc["nud"]))
    Lb(arch, bg, (ch.update_velocity_, (B["V"], B["rV"], B["Pr"], B["tau"], B["rho"], sc["eta_ve"], sc["nud"], bg)))
    check("update_velocity(field rho)")
    nd = len(n)
    rl = tuple(1 if d == nd - 1 else 0 for d in range(nd))
    inc_o = o.Inclusion(rl, (0.02,) * nd, 0.33, 1.0, 0.0)
    par = dict(zip(("x0", "y0", "z0"), (0.02,) * nd))
    inc_b = ch.FunctionField(ch.init_incl, bg, bloc(ch, rl), parameters={**par, "r": 0.33, "in": 1.0, "out": 0.0})
    o.launch(Lo, og, o.update_velocity, (O["V"], O["rV"], O["Pr"], O["tau"], inc_o, sc["eta_ve"], sc["nud"]))
    Lb(arch, bg, (ch.update_velocity_, (B["V"], B["rV"], B["Pr"], B["tau"], inc_b, sc["eta_ve"], sc["nud"], bg)))
    check("update_velocity(FunctionField rho)")
    o.launch(Lo, og, o.update_thermal_flux, (O["qT"], O["T"], O["V"], sc["lam"]))
    Lb(arch, bg, (ch.update_thermal_flux_, (B["qT"], B["T"], B["V"], sc["lam"], bg)))
    check("update_thermal_flux")
    o.launch(Lo, og, o.update_thermal, (O["T"], O["To"], O["qT"], sc["dt"]))
    Lb(arch, bg, (ch.update_thermal_, (B["T"], B["To"], B["qT"], sc["dt"], bg)))
    check("update_thermal")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [(64, 48), (255, 257)])
def test_diffusion_ops_bit_exact(ch, arch, oracle, n, dtype):
    o = oracle
    rng = np.random.default_rng(3)
    og, bg = mk_grids(ch, o, arch, n, dtype=dtype)
    Co, qo = o.Field(og, 0), o.VectorField(og)
    Cb, qb = ch.Field(arch, bg), ch.VectorField(arch, bg)
    for a, b in [(Co, Cb), (qo["x"], qb.x), (qo["y"], qb.y)]:
        fill_pair(rng, a, b)
    Lo, Lb = o.Launcher(og), ch.Launcher(arch, bg)
    chi, dt = dtype(1.7), dtype(1e-3)
    o.launch(Lo, og, o.compute_q, (qo, Co, chi))
    Lb(arch, bg, (ch.compute_q_, (qb, Cb, chi, bg)))
    o.launch(Lo, og, o.update_C, (Co, qo, dt))
    Lb(arch, bg, (ch.update_C_, (Cb, qb, dt, bg)))
    for nm, a, b in [("C", Co, Cb), ("qx", qo["x"], qb.x), ("qy", qo["y"], qb.y)]:
        assert_same(a, b, nm)


# ------------------------------------------------------------------------------------------------ whole solvers
def test_config1_diffusion_256(ch, arch, oracle):
    """BASELINE config 1: examples/diffusion_2d.jl, 256^2, outer_width (16, 8), nt = 100, C => Neumann()."""
    import drivers as OD
    from chmy_b200 import drivers as BD
    C0 = np.random.default_rng(0).random((256, 256))
    od = OD.Diffusion2D((256, 256), C0=C0)
    bd = BD.Diffusion2D(arch, (256, 256), C0=C0)
    od.run(100)
    bd.run(100)
    for k, f in od.fields().items():
        assert_same(f, bd.fields()[k], k)
    # the split launch (inner + 4 slabs on two streams) must give the same bits as the single launch
    bs = BD.Diffusion2D(arch, (256, 256), C0=C0, exact_split=True)
    bs.run(100)
    for k, f in od.fields().items():
        assert_same(f, bs.fields()[k], "split:" + k)


@pytest.mark.parametrize("n,fun", [((30, 22, 14), False), ((24, 24, 24), True), ((126, 126), True), ((63, 40), False)])
def test_stokes_solver_fields_and_residual_history(ch, arch, oracle, n, fun):
    """Configs 3/4 at oracle-sized grids: 2 outer steps x 110 PT iterations (thermal on in step 2), residual check
    every 22 iterations.  Full padded arrays and the whole residual history within 1e-12 relative."""
    import drivers as OD
    from chmy_b200 import drivers as BD
    osol = OD.Stokes(n, rho_g_function=fun)
    bsol = BD.Stokes(arch, n, rho_g_function=fun)
    ho = osol.run(2, 110, 22)
    hb = bsol.run(2, 110, 22)
    assert len(ho) == len(hb) == 10
    for a, b in zip(ho, hb):
        assert a[:2] == b[:2]
        for x, y in zip(a[2:], b[2:]):
            assert abs(x - y) <= 1e-12 * abs(x), (a, b)
    assert bsol.dt == osol.dt and bsol.eta_ve == osol.eta_ve
    bf = bsol.fields()
    for k, f in osol.fields().items():
        assert_same(f, bf[k], k, tol=1e-12)


# the parity grids BASELINE.md section 3 / SURVEY.md 8(d) name: odd sizes that straddle every tile edge of the tuned and fused
# kernels (60-cell row segments, 22-row clusters, 64-plane chunks), long/flat/tall aspect ratios, the 2D sizes
BASELINE_GRIDS = [(63, 63, 63), (127, 127, 127), (191, 129, 67), (255, 257), (1023, 1023)]
if __import__("os").environ.get("CHMY_DRYRUN") == "1":       # tests/test_gpu_suite_dryrun.py: same code, oracle-over-oracle sizes
    BASELINE_GRIDS = [(21, 19, 17), (40, 33)]


_ORACLE_RUNS = {}


@pytest.mark.parametrize("fusion", [3, 0], ids=["fused", "two-kernel"])
@pytest.mark.parametrize("n", BASELINE_GRIDS, ids=lambda n: "x".join(map(str, n)))
def test_baseline_parity_grids_fields_and_residual_history(ch, oracle, n, fusion):
    """SURVEY.md 8(d) parity runs: the Stokes driver (stokes_3d_inc_ve_T.jl / stokes_2d_inc_ve_T.jl) for 2 outer steps x 110 PT
    iterations (thermal sub-steps in the second), residual check every 22 iterations -- every field's FULL padded array and
    the whole residual history within 1e-12 relative of the oracle, with the fused sweeps (the bench path: batches folded,
    frame carry-over elided) and with the two tuned kernels."""
    import drivers as OD
    from chmy_b200 import drivers as BD
    a = ch.Arch(ch.B200Backend())
    try:
        ch.set_fusion(a, fusion)
        if n not in _ORACLE_RUNS:                                 # one oracle run serves both parametrisations
            osol = OD.Stokes(n, rho_g_function=True)
            _ORACLE_RUNS.clear()
            _ORACLE_RUNS[n] = (osol, osol.run(2, 110, 22))
        osol, ho = _ORACLE_RUNS[n]
        bsol = BD.Stokes(a, n, rho_g_function=True, blocking=False)
        hb = bsol.run(2, 110, 22)
        assert len(ho) == len(hb) == 10
        for x, y in zip(ho, hb):
            assert x[:2] == y[:2]
            for p, q in zip(x[2:], y[2:]):
                assert abs(p - q) <= 1e-12 * abs(p), (x, y)
        assert bsol.dt == osol.dt and bsol.eta_ve == osol.eta_ve
        if fusion:
            assert ch.fused_count(a) == 220 + 110            # every mechanics pair and every thermal pair ran as a sweep
        bf = bsol.fields()
        for k, f in osol.fields().items():
            assert_same(f, bf[k], k, tol=1e-12)
    finally:
        a.close()


def test_stokes_split_equals_unsplit(ch, arch):
    """Launcher with outer_width honoured literally (two streams) vs one full-range kernel: identical bits."""
    from chmy_b200 import drivers as BD
    a = BD.Stokes(arch, (40, 36, 28), rho_g_function=True)
    b = BD.Stokes(arch, (40, 36, 28), rho_g_function=True, outer_width=(8, 4, 3), exact_split=True, blocking=False)
    a.run(2, 40, 20)
    b.run(2, 40, 20)
    assert a.history == b.history
    fb = b.fields()
    for k, f in a.fields().items():
        assert np.array_equal(f.parent(), fb[k].parent()), k


# ------------------------------------------------------------------------------------------------ tuned kernels
def _set_tuning(disable_fast=-1, true_div=-1):
    from chmy_b200 import _lib as L
    L.check(L.lib().chmy_set_tuning(disable_fast, true_div))


@pytest.mark.parametrize("c", [3.0, 10.0, 0.0171 * 1.3, 0.737, 1.0 / 3.0, 6.02e23, 1.7e-19, 1.9999999999999998])
def test_exact_division_by_uniform_scalar(ch, arch, c):
    """The exact-division sequence (double-double reciprocal product + one Markstein correction) must equal IEEE division bit for bit (2^28 operands per divisor:
    random significands over 120 binades, exact multiples of c and their 1-ulp neighbours)."""
    import ctypes as C
    from chmy_b200 import _lib as L
    bad, used = C.c_ulonglong(1), C.c_int(-1)
    L.check(L.lib().chmy_selftest_division(arch.ctx, c, 1 << 28, 12345, C.byref(bad), C.byref(used)))
    if c == 1.9999999999999998:
        assert used.value == 0                       # all-ones significand -> routed to true division
    else:
        assert used.value == 1 and bad.value == 0, (c, bad.value)


@pytest.mark.parametrize("c", [3.0, 10.0, 0.0171 * 1.3, 0.737, 1.0 / 3.0, 6.02e23, 1.7e-19, 3.8780364835615564, 0.007735422636398862])
def test_two_operation_division_where_it_is_proven(ch, arch, c):
    """the fused 3D sweep divides in two operations when div2_exact proves that exact for the divisor (include/chmy_b200.h:
    chmy_division_two_op_exact); on the device the sequence then equals IEEE division on 2^28 operands, and divisors with a
    failing operand (the last two: found by the number-theoretic search) are refused"""
    import ctypes as C
    from chmy_b200 import _lib as L
    bad, proved = C.c_ulonglong(1), C.c_int(-1)
    L.check(L.lib().chmy_selftest_division2(arch.ctx, c, 1 << 28, 54321, C.byref(bad), C.byref(proved)))
    assert proved.value == int(ch.division_two_op_exact(c))
    if c in (3.8780364835615564, 0.007735422636398862):
        assert proved.value == 0
    else:
        assert proved.value == 1 and bad.value == 0, (c, bad.value)


@pytest.mark.parametrize("n", [(130, 19, 70), (63, 9, 5), (64, 8, 64), (65, 17, 65), (1, 1, 1),
                               (300, 70), (257, 129), (256, 64), (255, 63), (1, 1), (515, 3)])
@pytest.mark.parametrize("true_div", [0, 1])
def test_fast_kernels_equal_generic_kernels(ch, arch, n, true_div):
    """Marching / vectorised kernels (3D: ops_fast.cu, 2D: ops_fast2d.cu; stress, velocity, thermal pair) vs the
    one-thread-per-cell kernels: identical bits on the full padded arrays for tile-edge sizes (x not a multiple of
    64 or 256, odd y/z, single cells)."""
    from chmy_b200 import drivers as BD
    nd = len(n)
    res = []
    for fast in (1, 0):
        _set_tuning(disable_fast=0 if fast else 1, true_div=true_div)
        s = BD.Stokes(arch, n, rho_g_function=(n[0] % 2 == 1))
        rng = np.random.default_rng(42)
        for f in s.fields().values():
            f.from_host(rng.random(tuple(d + 4 for d in f.dims)) - 0.5, [-1] * nd, [d + 2 for d in f.dims])
        s.begin_time_step()
        for _ in range(3):
            s.mechanics()
            s.thermal()
        res.append({k: f.parent() for k, f in s.fields().items()})
    _set_tuning(0, 0)
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k


@pytest.mark.parametrize("n", [(300, 70), (257, 129), (64, 64), (3, 2)])
def test_fast_diffusion_equals_generic(ch, arch, n):
    """2D diffusion: tuned y-marching kernels vs generic kernels, with and without the split launch."""
    from chmy_b200 import drivers as BD
    C0 = np.random.default_rng(5).random(n)
    res = []
    for fast, split in ((1, False), (0, False), (1, True)):
        _set_tuning(disable_fast=0 if fast else 1, true_div=0)
        ow = (16, 8) if min(n) >= 32 else None
        s = BD.Diffusion2D(arch, n, C0=C0, outer_width=ow, exact_split=split and ow is not None)
        s.run(7)
        res.append({k: f.parent() for k, f in s.fields().items()})
    _set_tuning(0, 0)
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
        assert np.array_equal(res[0][k], res[2][k]), "split:" + k
