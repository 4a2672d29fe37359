"""
GPU tests of the entry points added late in round 1 (operators, Float32 pieces, pinned arrays, batched maxima).
Same bar as tests/test_b200_parity.py: through the C ABI, against the CPU oracle, bit-exact unless stated.
"""
import numpy as np
import pytest

from helpers import assert_same, fill_pair

# Green on the driver's B200 at the end of round 1 (GPUTEST_r01.json: all XPASS) -- enforced since round 2.
pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    return chmy_b200


@pytest.fixture(scope="module")
def arch(ch):
    a = ch.Arch(ch.B200Backend())
    yield a
    a.close()


def mk_grids(ch, o, arch, n):
    nd = len(n)
    origin = tuple(-1.0 - 0.1 * d for d in range(nd))
    extent = tuple(2.0 + 0.3 * d for d in range(nd))
    return o.Grid(origin, extent, n), ch.UniformGrid(arch, origin=origin, extent=extent, dims=n)


# ------------------------------------------------------------------------------------------------ pinned host arrays
@pytest.mark.parametrize("n,loc", [((33, 18), (0, 1)), ((12, 10, 8), (1, 0, 1))])
def test_pinned_host_array_roundtrip(ch, arch, oracle, n, loc):
    """set!(f, A) / Array(interior(f)) with A in page-locked memory (chmy_host_alloc): same bits as pageable memory."""
    og, bg = mk_grids(ch, oracle, arch, n)
    of = oracle.Field(og, loc)
    bf = ch.Field(arch, bg, tuple(ch.Vertex() if l else ch.Center() for l in loc))
    rng = np.random.default_rng(11)
    a = ch.pinned_array(arch, bf.dims)
    assert a.shape == tuple(bf.dims) and a.flags.f_contiguous and a.flags.writeable
    a[...] = rng.random(bf.dims)
    of.set(a)
    ch.set_(bf, a)
    assert_same(of, bf, "set!(f, pinned A)")
    out = ch.pinned_array(arch, bf.dims)
    got = ch.interior(bf, out=out)
    assert got is out and np.array_equal(out, a)
    with pytest.raises(ValueError):
        ch.interior(bf, out=np.zeros(tuple(d + 1 for d in bf.dims), order="F"))
    del a, out, got          # frees the pinned buffers (finalizer -> chmy_host_free)


# ------------------------------------------------------------------------------------------------ grid operators
import itertools


def flip(loc, d):
    return tuple(1 - l if a == d else l for a, l in enumerate(loc))


def bloc(ch, loc):
    return tuple(ch.Vertex() if l else ch.Center() for l in loc)


def pair(ch, oracle, arch, og, bg, loc, rng, positive=False, layout=0):
    of = oracle.Field(og, loc)
    bf = ch.Field(arch, bg, bloc(ch, loc), layout=layout)
    a = (rng.random(of.sdims) + 0.5) if positive else (rng.random(of.sdims) - 0.5)
    of.data[...] = a                                                   # interior, halo AND padding
    bf.from_host(a, [-1] * len(of.dims), [d + 2 for d in of.dims])
    return of, bf


KOP = {"left": "left_", "right": "right_", "delta": "delta_", "partial": "partial_", "partial2": "partial2_", "dkd": "dkd_"}


def both(ch, oracle, arch, og, bg, kind, dst, src, k=None, dim=0):
    """run `kind` through the oracle and through launch(arch, grid, op => (dst, src..., grid)); compare full padded arrays"""
    dst = dst if isinstance(dst, list) else [dst]
    src = src if isinstance(src, list) else [src]
    oracle.apply_operator(og, kind, [d[0] for d in dst], [s[0] for s in src], k=None if k is None else k[0], dim=dim)
    op = getattr(ch, KOP[kind])(dim + 1) if kind in KOP else getattr(ch, kind + "_")
    bd = [d[1] for d in dst]
    bs = [s[1] for s in src]
    args = (bd[0] if len(bd) == 1 else bd, bs[0] if len(bs) == 1 else bs) + (() if k is None else (k[1],)) + (bg,)
    ch.Launcher(arch, bg)(arch, bg, (op, args))
    for q, (of, bf) in enumerate(dst):
        assert_same(of, bf, f"{kind} dim={dim} dst{q}")
    for of, bf in src:
        assert_same(of, bf, f"{kind}: source modified")


@pytest.mark.parametrize("n,layout", [((9,), 0), ((7, 5), 0), ((7, 5), 1), ((6, 5, 4), 0), ((70, 9, 5), 0)])
def test_grid_operators_every_location_bit_exact(ch, arch, oracle, n, layout):
    nd = len(n)
    og, bg = mk_grids(ch, oracle, arch, n)
    rng = np.random.default_rng(5)
    mk = lambda loc, positive=False: pair(ch, oracle, arch, og, bg, loc, rng, positive, layout)
    locs = list(itertools.product((0, 1), repeat=nd))
    for loc in locs:
        f = mk(loc)
        for dim in range(nd):
            for kind in ("left", "right", "delta", "partial"):
                both(ch, oracle, arch, og, bg, kind, mk(flip(loc, dim)), f, dim=dim)
            both(ch, oracle, arch, og, bg, "partial2", mk(loc), f, dim=dim)
            for kloc in (locs[0], locs[-1], locs[len(locs) // 2]):
                both(ch, oracle, arch, og, bg, "dkd", mk(loc), f, k=mk(kloc), dim=dim)
        both(ch, oracle, arch, og, bg, "lapl", mk(loc), f)
        for kloc in locs:
            both(ch, oracle, arch, og, bg, "divg_grad", mk(loc), f, k=mk(kloc))
        fpos = mk(loc, True)
        for to in locs:
            both(ch, oracle, arch, og, bg, "lerp", mk(to), f)
            both(ch, oracle, arch, og, bg, "hlerp", mk(to), fpos)
    ctr, vtx = (0,) * nd, (1,) * nd
    V = [mk(flip(ctr, d)) for d in range(nd)]
    both(ch, oracle, arch, og, bg, "divg", mk(ctr), V)
    both(ch, oracle, arch, og, bg, "vmag", mk(ctr), V)
    for floc in (ctr, vtx):
        f = mk(floc)
        both(ch, oracle, arch, og, bg, "grad", [mk(flip(floc, d)) for d in range(nd)], f)
        both(ch, oracle, arch, og, bg, "kgrad", [mk(flip(floc, d)) for d in range(nd)], f, k=mk(locs[-1]))


def test_reference_operator_identities(ch, arch):
    """test/test_grid_operators.jl:13-131 on the B200 path: divg == sum of partials (`==`, :41), lapl == sum of second
    derivatives (`==`, :61), divg_grad ≈ divg(lerp(χ) grad) for χ at Center and Vertex (:98,:110), vmag of (2,2,2) (:130)."""
    grid = ch.UniformGrid(arch, origin=(-5.0, -5.0, -5.0), extent=(10.0, 10.0, 10.0), dims=(12, 10, 8))
    launch = ch.Launcher(arch, grid)
    gauss = lambda x, y, z: np.exp(-x ** 2 - y ** 2 - z ** 2)
    F = lambda loc=None: ch.Field(arch, grid, loc if loc is not None else ch.Center())
    Ci, C1, C2, P = F(), F(), F(), F()
    V = ch.VectorField(arch, grid)
    ch.set_(Ci, grid, gauss)
    launch(arch, grid, (ch.grad_, (V, Ci, grid)))                                   # divg1!
    acc = None
    for d, c in enumerate(V):
        launch(arch, grid, (ch.partial_(d + 1), (P, c, grid)))
        acc = ch.interior(P) if acc is None else acc + ch.interior(P)
    launch(arch, grid, (ch.divg_, (C2, V, grid)))
    assert np.array_equal(ch.interior(C2), acc) and np.abs(acc).max() > 1e-3       # :41
    acc = None
    for d in range(3):
        launch(arch, grid, (ch.partial2_(d + 1), (P, Ci, grid)))
        acc = ch.interior(P) if acc is None else acc + ch.interior(P)
    launch(arch, grid, (ch.lapl_, (C2, Ci, grid)))
    assert np.array_equal(ch.interior(C2), acc)                                     # :61
    for chi in (F(), F(ch.Vertex())):                                               # :86-111
        ch.set_(chi, grid, gauss)
        ch.set_(C1, 0.0); ch.set_(C2, 0.0)
        launch(arch, grid, (ch.kgrad_, (V, Ci, chi, grid)))                         # divg_grad1!
        launch(arch, grid, (ch.divg_, (C1, V, grid)))
        launch(arch, grid, (ch.divg_grad_, (C2, Ci, chi, grid)))
        assert np.allclose(ch.interior(C2), ch.interior(C1), rtol=1.5e-8, atol=0.0) or \
            np.allclose(ch.interior(C2), ch.interior(C1), rtol=1.5e-8, atol=1e-14)
    for c in V:
        ch.set_(c, 2.0)
    launch(arch, grid, (ch.vmag_, (C1, V, grid)))
    v = ch.interior(C1)
    assert np.all(np.vectorize(lambda x: float(f"{x:.5g}"))(v) == 3.4641)           # :130


def test_reference_interpolations(ch, arch):
    """test/test_interpolations.jl:15-74 through lerp_: c2v, c2c, c2cv, c2vc, v2c, v2v, v2cv, v2vc on the 2x2 grid."""
    grid = ch.UniformGrid(arch, origin=(0.0, 0.0), extent=(1.0, 1.0), dims=(2, 2))
    launch = ch.Launcher(arch, grid)
    av4 = lambda A: 0.25 * (A[:-1, :-1] + A[1:, :-1] + A[1:, 1:] + A[:-1, 1:])
    avx = lambda A: 0.5 * (A[:-1, :] + A[1:, :])
    avy = lambda A: 0.5 * (A[:, :-1] + A[:, 1:])
    Cn, Vx = ch.Center(), ch.Vertex()

    def interp(src, to):
        dst = ch.Field(arch, grid, to)
        launch(arch, grid, (ch.lerp_, (dst, src, grid)))
        return ch.interior(dst)

    fc = ch.Field(arch, grid, Cn)
    ch.set_(fc, np.arange(1, 5, dtype=float).reshape((2, 2), order="F"))
    fci = ch.interior(fc)
    assert np.allclose(interp(fc, Vx)[1:-1, 1:-1], av4(fci))
    assert np.allclose(interp(fc, Cn), fci)
    assert np.allclose(interp(fc, (Cn, Vx))[:, 1:-1], avy(fci))
    assert np.allclose(interp(fc, (Vx, Cn))[1:-1, :], avx(fci))
    fv = ch.Field(arch, grid, Vx)
    ch.set_(fv, np.arange(1, 10, dtype=float).reshape((3, 3), order="F"))
    fvi = ch.interior(fv)
    assert np.allclose(interp(fv, Cn), av4(fvi))
    assert np.allclose(interp(fv, Vx), fvi)
    assert np.allclose(interp(fv, (Cn, Vx)), avx(fvi))
    assert np.allclose(interp(fv, (Vx, Cn)), avy(fvi))


def test_operator_location_checks(ch, arch):
    """the library refuses operands that do not sit where the reference's operator reads / produces them"""
    grid = ch.UniformGrid(arch, origin=(0.0, 0.0), extent=(1.0, 1.0), dims=(6, 5))
    launch = ch.Launcher(arch, grid)
    c, c2, v = ch.Field(arch, grid, ch.Center()), ch.Field(arch, grid, ch.Center()), ch.Field(arch, grid, ch.Vertex())
    with pytest.raises(ch.ChmyError):
        launch(arch, grid, (ch.partial_(1), (c2, c, grid)))          # ∂x of a Center field lives at (Vertex, Center)
    with pytest.raises(ch.ChmyError):
        launch(arch, grid, (ch.lapl_, (v, c, grid)))
    with pytest.raises(ch.ChmyError):
        launch(arch, grid, (ch.lerp_, (c, c, grid)))                 # dst aliases the source
    with pytest.raises(ch.ChmyError):
        launch(arch, grid, (ch.partial_(3), (c2, c, grid)))          # Dim(3) on a 2D grid
    with pytest.raises(ValueError):
        ch.partial_(0)


# ------------------------------------------------------------------------------------------------ Float32 instantiation
# The reference's tests run for T in (Float32, Float64) (test/common.jl:9).  Fields, set!/copies, maxabs, bc!, halo
# slabs and the grid operators exist in Float32 on this path; the solver ops are Float64 programs.
F32 = np.float32


def mk_grids32(ch, o, arch, n, origin=None, extent=None):
    nd = len(n)
    origin = origin or tuple(-1.0 - 0.1 * d for d in range(nd))
    extent = extent or tuple(2.0 + 0.3 * d for d in range(nd))
    return o.Grid(origin, extent, n, dtype=F32), ch.UniformGrid(arch, origin=origin, extent=extent, dims=n, dtype=F32)


def fill_pair32(rng, of, bf, scale=1.0):
    a = ((rng.random(of.sdims) - 0.5) * scale).astype(F32)
    of.data[...] = a
    bf.from_host(a, [-1] * len(of.dims), [d + 2 for d in of.dims])
    return a


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,loc", [((7,), (1,)), ((33, 18), (0, 1)), ((12, 10, 8), (1, 0, 1)), ((70, 5, 3), (0, 0, 0))])
def test_f32_field_roundtrip_fill_copy_maxabs(ch, arch, oracle, n, loc, layout):
    og, bg = mk_grids32(ch, oracle, arch, n)
    for d in range(len(n)):                                              # host grid numbers == oracle, in binary32
        assert float(ch.spacing(bg)[d]) == og.spacing[d] and float(ch.inv_spacing(bg)[d]) == og.inv_spacing[d]
    of = oracle.Field(og, loc)
    bf = ch.Field(arch, bg, bloc(ch, loc), layout=layout)
    assert bf.dtype == F32 and bf.dims == of.dims and bf.info().dtype == 1
    assert np.array_equal(bf.parent(), np.zeros(of.sdims, F32)) and bf.parent().dtype == F32
    if layout == 0:
        assert (bf.info().origin_ptr - 4) % 128 == 0                     # logical index 0 of each row is 128 B aligned
    rng = np.random.default_rng(21)
    a = fill_pair32(rng, of, bf)
    assert_same(of, bf, "roundtrip f32")
    assert np.array_equal(ch.interior(bf, with_halo=True), of.interior(with_halo=True))
    assert ch.maxabs(bf) == of.maxabs() == float(np.abs(of.interior()).max())
    ch.set_(bf, 3.5); of.set(3.5)
    assert_same(of, bf, "set scalar f32")
    A = rng.random(of.dims).astype(F32)
    ch.set_(bf, A); of.set(A)
    assert_same(of, bf, "set array f32")
    bf2 = ch.Field(arch, bg, bloc(ch, loc), layout=layout)
    ch.set_(bf2, bf)
    assert np.array_equal(ch.interior(bf2), of.interior()) and np.array_equal(bf2.parent()[0], np.zeros_like(bf2.parent()[0]))
    a[tuple(3 for _ in n)] = np.nan
    bf.from_host(a, [-1] * len(n), [d + 2 for d in of.dims])
    assert np.isnan(ch.maxabs(bf))
    f64 = ch.Field(arch, ch.UniformGrid(arch, origin=(0.0,) * len(n), extent=(1.0,) * len(n), dims=n), bloc(ch, loc))
    with pytest.raises(ch.ChmyError):
        ch.set_(f64, bf)                                                  # element types differ


@pytest.mark.parametrize("n", [(40, 27), (19, 14, 11)])
def test_f32_set_inclusion(ch, arch, oracle, n):
    nd = len(n)
    og, bg = mk_grids32(ch, oracle, arch, n, origin=(-1.0,) * nd, extent=(2.0,) * nd)
    for loc in [(0,) * nd, tuple(1 if d == nd - 1 else 0 for d in range(nd))]:
        of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
        oracle.set_inclusion(of, oracle.Inclusion(loc, (0.05,) * nd, 0.31, 1.0, 0.1))
        par = dict(zip(("x0", "y0", "z0"), (0.05,) * nd))
        ch.set_(bf, bg, ch.init_incl, parameters={**par, "r": 0.31, "in": 1.0, "out": 0.1})
        assert_same(of, bf, "inclusion f32")
        assert 0 < (of.interior() == 1.0).sum() < of.interior().size


@pytest.mark.parametrize("n,loc", [((8,), (0,)), ((8,), (1,)), ((8, 8), (0, 1)), ((8, 8, 6), (0, 1, 0)), ((13, 7, 5), (1, 1, 0))])
def test_f32_bc_matches_oracle_bit_exact(ch, arch, oracle, n, loc):
    """test/test_boundary_conditions.jl for T = Float32: every touched cell equals the binary32 oracle."""
    import math
    nd = len(n)
    og, bg = mk_grids32(ch, oracle, arch, n, origin=(-math.pi,) * nd, extent=(2 * math.pi,) * nd)
    of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    rng = np.random.default_rng(7)
    for mk_o, mk_b in [(oracle.Dirichlet, ch.Dirichlet), (oracle.Neumann, ch.Neumann)]:
        for val in (None, 2.0, 0.3):
            fill_pair32(rng, of, bf)
            oracle.bc_(og, (of, mk_o(val)))
            ch.bc_(arch, bg, (bf, mk_b(val)))
            assert_same(of, bf, f"bc f32 {mk_o.__name__}({val})")
    if nd >= 2:
        D = 0
        tg_o = oracle.transverse_grid(og, D)
        keep = [a for a in range(nd) if a != D]
        tg_b = ch.UniformGrid(arch, origin=[og.origin[a] for a in keep], extent=[og.extent[a] for a in keep],
                              dims=[n[a] for a in keep], dtype=F32)
        vo, vb = oracle.Field(tg_o, oracle.VERTEX), ch.Field(arch, tg_b, ch.Vertex())
        fill_pair32(rng, vo, vb)
        fill_pair32(rng, of, bf)
        oracle.bc_(og, (of, {"x": oracle.Neumann(vo)}))
        ch.bc_(arch, bg, (bf, {"x": ch.Neumann(vb)}))
        assert_same(of, bf, "field-valued Neumann f32")


@pytest.mark.parametrize("n,loc", [((9, 6), (1, 0)), ((11, 7, 5), (0, 1, 1))])
def test_f32_halo_pack_unpack_bit_exact(ch, arch, oracle, n, loc):
    import ctypes as C
    from chmy_b200 import _lib as L
    og, bg = mk_grids32(ch, oracle, arch, n)
    of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    enc = (np.arange(of.data.size, dtype=np.float64).reshape(of.sdims, order="F") + 1e3).astype(F32)
    of.data[...] = enc
    bf.from_host(enc, [-1] * len(n), [d + 2 for d in of.dims])
    for D in range(len(n)):
        for S in range(2):
            ref = oracle.pack_send(of, D, S)
            buf = np.empty(ref.size, dtype=F32)
            L.check(L.lib().chmy_halo_pack(arch.ctx, bf.handle, D, S, buf.ctypes.data_as(C.c_void_p)))
            assert ref.dtype == F32 and np.array_equal(buf, ref), (D, S)
            msg = (-ref[::-1]).copy()
            oracle.unpack_recv(of, D, S, msg)
            L.check(L.lib().chmy_halo_unpack(arch.ctx, bf.handle, D, S, msg.ctypes.data_as(C.c_void_p)))
            assert_same(of, bf, f"unpack f32 {D},{S}")


@pytest.mark.parametrize("n", [(9,), (7, 5), (6, 5, 4)])
def test_f32_grid_operators_bit_exact(ch, arch, oracle, n):
    nd = len(n)
    og, bg = mk_grids32(ch, oracle, arch, n)
    rng = np.random.default_rng(6)

    def mk(loc, positive=False):
        of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
        a = (rng.random(of.sdims) + (0.5 if positive else -0.5)).astype(F32)
        of.data[...] = a
        bf.from_host(a, [-1] * nd, [d + 2 for d in of.dims])
        return of, bf
    locs = list(itertools.product((0, 1), repeat=nd))
    for loc in locs:
        f = mk(loc)
        for dim in range(nd):
            both(ch, oracle, arch, og, bg, "partial", mk(flip(loc, dim)), f, dim=dim)
            both(ch, oracle, arch, og, bg, "partial2", mk(loc), f, dim=dim)
            both(ch, oracle, arch, og, bg, "dkd", mk(loc), f, k=mk(locs[-1]), dim=dim)
        both(ch, oracle, arch, og, bg, "lapl", mk(loc), f)
        both(ch, oracle, arch, og, bg, "divg_grad", mk(loc), f, k=mk(locs[0]))
        fpos = mk(loc, True)
        for to in locs:
            both(ch, oracle, arch, og, bg, "lerp", mk(to), f)
            both(ch, oracle, arch, og, bg, "hlerp", mk(to), fpos)
    ctr = (0,) * nd
    V = [mk(flip(ctr, d)) for d in range(nd)]
    both(ch, oracle, arch, og, bg, "divg", mk(ctr), V)
    both(ch, oracle, arch, og, bg, "vmag", mk(ctr), V)
    both(ch, oracle, arch, og, bg, "kgrad", [mk(flip(ctr, d)) for d in range(nd)], mk(ctr), k=mk(locs[-1]))


def test_one_launch_is_of_one_element_type(ch, arch):
    """the kernels are generic in the element type (tests/test_b200_parity.py runs every op in both); mixing the two in one
    launch is an error, not a conversion"""
    g = ch.UniformGrid(arch, origin=(0.0, 0.0), extent=(1.0, 1.0), dims=(8, 6), dtype=F32)
    C, q = ch.Field(arch, g, ch.Center()), ch.VectorField(arch, g)
    assert C.dtype == F32 and q.x.dtype == F32
    ch.Launcher(arch, g)(arch, g, (ch.compute_q_, (q, C, F32(1.0), g)))
    f64 = ch.Field(arch, ch.UniformGrid(arch, origin=(0.0, 0.0), extent=(1.0, 1.0), dims=(8, 6)), ch.Center())
    with pytest.raises(ch.ChmyError):
        ch.Launcher(arch, g)(arch, g, (ch.compute_q_, (q, f64, F32(1.0), g)))
    with pytest.raises(ch.ChmyError):
        ch.Launcher(arch, g)(arch, g, (ch.lapl_, (C, f64, g)))           # operators need one element type


# ------------------------------------------------------------------------------------------------ FunctionField, any body
@pytest.mark.parametrize("n", [(24, 18), (14, 10, 8)])
def test_function_field_with_an_arbitrary_body(ch, arch, oracle, n):
    """function_field.jl:49-59 with a body other than the drivers' init_incl: update_velocity! with
    rho_g = FunctionField(f, grid, loc; parameters) must equal the oracle run with a stored field holding f at the
    coordinates of every index the kernel reads (the reference evaluates f in-kernel at those coordinates)."""
    o, nd = oracle, len(n)
    og, bg = mk_grids(ch, o, arch, n)
    rng = np.random.default_rng(12)
    rl = tuple(1 if d == nd - 1 else 0 for d in range(nd))
    body = (lambda x, y, a: a * x - y * y) if nd == 2 else (lambda x, y, z, a: a * x - y * y + 0.5 * z)
    ff = ch.FunctionField(body, bg, bloc(ch, rl), parameters=(0.75,))
    rho_o = o.Field(og, rl)
    rho_o.data[tuple(slice(1, -1) for _ in n)] = ff.values()                  # logical 0..d+1
    Vo, rVo, tauo, Pro = o.VectorField(og), o.VectorField(og), o.TensorField(og), o.Field(og, 0)
    Vb, rVb, taub, Prb = ch.VectorField(arch, bg), ch.VectorField(arch, bg), ch.TensorField(arch, bg), ch.Field(arch, bg)
    for fo, fb in list(zip(Vo.values(), Vb)) + list(zip(rVo.values(), rVb)) + list(zip(tauo.values(), taub)) + [(Pro, Prb)]:
        fill_pair(rng, fo, fb)
    o.launch(o.Launcher(og), og, o.update_velocity, (Vo, rVo, Pro, tauo, rho_o, 0.737, 0.00931))
    ch.Launcher(arch, bg)(arch, bg, (ch.update_velocity_, (Vb, rVb, Prb, taub, ff, 0.737, 0.00931, bg)))
    for fo, fb in list(zip(Vo.values(), Vb)) + list(zip(rVo.values(), rVb)):
        assert_same(fo, fb, "update_velocity!(FunctionField with an arbitrary body)")


# ------------------------------------------------------------------------------------------------ 1D halo slabs
@pytest.mark.parametrize("loc", [(0,), (1,)])
def test_halo_pack_unpack_1d(ch, arch, oracle, loc):
    """communication_views.jl:1-34 on a 1D field: the slab is a single element (send index 1+overlap | d-overlap, recv
    index 0 | d+1)."""
    import ctypes as C
    from chmy_b200 import _lib as L
    og, bg = mk_grids(ch, oracle, arch, (9,))
    of, bf = oracle.Field(og, loc), ch.Field(arch, bg, bloc(ch, loc))
    enc = np.arange(of.data.size, dtype=np.float64) + 1e3
    of.data[...] = enc
    bf.from_host(enc, [-1], [of.dims[0] + 2])
    for S in range(2):
        ref = oracle.pack_send(of, 0, S)
        assert ref.size == 1
        buf = np.empty(1)
        L.check(L.lib().chmy_halo_pack(arch.ctx, bf.handle, 0, S, buf.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(buf, ref)
        msg = -ref
        oracle.unpack_recv(of, 0, S, msg)
        L.check(L.lib().chmy_halo_unpack(arch.ctx, bf.handle, 0, S, msg.ctypes.data_as(C.c_void_p)))
        assert_same(of, bf, f"1D unpack side {S}")


# ------------------------------------------------------------------------------------------------ set!(...; discrete=true)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_set_discrete_known_answers(ch, arch, dtype):
    """test/test_fields.jl:21-38 (discrete bodies get (grid, loc, ix, iy, iz, params...)) for both element types, `==`."""
    g = ch.UniformGrid(arch, origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0), dims=(2, 2, 2), dtype=dtype)
    f = ch.Field(arch, g, (ch.Center(), ch.Vertex(), ch.Center()))
    exp_y = np.zeros((2, 3, 2), dtype); exp_y[:, 1, :], exp_y[:, 2, :] = 0.5, 1.0
    exp_x = np.zeros((2, 3, 2), dtype); exp_x[0], exp_x[1] = 0.25, 0.75
    ch.fill_parent_(f, float("nan"))
    ch.set_(f, g, lambda grid, loc, ix, iy, iz: ch.coord(grid, loc, 2, iy), discrete=True)
    assert np.array_equal(ch.interior(f), exp_y) and ch.interior(f).dtype == dtype
    ch.set_(f, g, lambda grid, loc, ix, iy, iz: ch.coord(grid, loc, 1, ix), discrete=True)
    assert np.array_equal(ch.interior(f), exp_x)
    ch.set_(f, g, lambda grid, loc, ix, iy, iz, sc: ch.coord(grid, loc, 2, iy) * sc, discrete=True, parameters=(dtype(2.0),))
    assert np.array_equal(ch.interior(f), 2 * exp_y)
    assert np.isnan(f.parent()[0]).all()                                # halo / padding untouched


@pytest.mark.parametrize("n", [(24, 18), (14, 10, 8), (70, 20, 9)])
def test_constant_field_body_force(ch, arch, oracle, n):
    """rho_g = ValueField(c) (src/Fields/constant_field.jl): update_velocity! must equal the oracle run with a stored field
    that holds c everywhere (tuned and generic kernels, FunctionField code path with coinciding in/out values)."""
    o, nd = oracle, len(n)
    og, bg = mk_grids(ch, o, arch, n)
    rng = np.random.default_rng(13)
    rl = tuple(1 if d == nd - 1 else 0 for d in range(nd))
    rho_o = o.Field(og, rl)
    rho_o.data[...] = 0.37
    Vo, rVo, tauo, Pro = o.VectorField(og), o.VectorField(og), o.TensorField(og), o.Field(og, 0)
    Vb, rVb, taub, Prb = ch.VectorField(arch, bg), ch.VectorField(arch, bg), ch.TensorField(arch, bg), ch.Field(arch, bg)
    for fo, fb in list(zip(Vo.values(), Vb)) + list(zip(rVo.values(), rVb)) + list(zip(tauo.values(), taub)) + [(Pro, Prb)]:
        fill_pair(rng, fo, fb)
    o.launch(o.Launcher(og), og, o.update_velocity, (Vo, rVo, Pro, tauo, rho_o, 0.737, 0.00931))
    ch.Launcher(arch, bg)(arch, bg, (ch.update_velocity_, (Vb, rVb, Prb, taub, ch.ValueField(0.37), 0.737, 0.00931, bg)))
    for fo, fb in list(zip(Vo.values(), Vb)) + list(zip(rVo.values(), rVb)):
        assert_same(fo, fb, "update_velocity!(ValueField)")


def test_architectures_known_answers(ch):
    """test/test_architectures.jl:5-27 on the B200 backend."""
    backend = ch.B200Backend()
    arch = ch.SingleDeviceArchitecture(backend, 1)
    try:
        assert isinstance(arch, ch.SingleDeviceArchitecture) and arch.backend is backend and arch.device == 1
        arch2 = ch.SingleDeviceArchitecture(arch)
        try:
            assert arch2.backend is arch.backend and arch2.device == arch.device
        finally:
            arch2.close()
        assert ch.get_backend(arch) is backend and ch.get_device(arch) == 1 and ch.set_device_(arch.device) == 1
        assert ch.is_gpu_aware(arch) is True
    finally:
        arch.close()


# ------------------------------------------------------------------------------------------------ batched residual maxima
@pytest.mark.parametrize("n", [(33, 18), (12, 10, 8), (9,)])
def test_maxabs_many_equals_the_single_reductions_and_the_oracle(ch, arch, oracle, n):
    """chmy_field_maxabs_many: the residual check's maxima (stokes_3d_inc_ve_T.jl:171-175) in one round trip -- same values
    as one chmy_field_maxabs per field and as the oracle (max is order-independent: exact), mixed locations and element
    types, NaN propagated like Julia's maximum."""
    og, bg = mk_grids(ch, oracle, arch, n)
    nd = len(n)
    rng = np.random.default_rng(21)
    ofs, bfs = [], []
    for q, loc in enumerate(itertools.islice(itertools.cycle(itertools.product((0, 1), repeat=nd)), 5)):
        of = oracle.Field(og, loc)
        bf = ch.Field(arch, bg, tuple(ch.Vertex() if l else ch.Center() for l in loc))
        fill_pair(rng, of, bf, scale=10.0 ** (q - 2))
        ofs.append(of)
        bfs.append(bf)
    many = ch.maxabs_many(*bfs)
    assert len(many) == 5
    for of, bf, m in zip(ofs, bfs, many):
        assert m == ch.maxabs(bf) == of.maxabs() == float(np.abs(of.interior()).max())
    halo = ch.maxabs_many(*bfs, with_halo=True)
    for of, m in zip(ofs, halo):
        assert m == float(np.abs(of.interior(with_halo=True)).max())
    # a NaN in the interior of one field shows in that field's maximum only
    bad = ofs[2].data.copy()
    bad[tuple(3 for _ in range(nd))] = np.nan
    bfs[2].from_host(bad, [-1] * nd, [d + 2 for d in ofs[2].dims])
    many2 = ch.maxabs_many(*bfs)
    assert np.isnan(many2[2]) and many2[:2] == many[:2] and many2[3:] == many[3:]
    assert ch.maxabs_many() == ()
    # Float32 fields ride in the same call
    bf32 = ch.Field(arch, ch.UniformGrid(arch, origin=(0.0,) * nd, extent=(1.0,) * nd, dims=n, dtype=np.float32), ch.Center())
    a32 = (rng.random(bf32.dims) - 0.5).astype(np.float32)
    ch.set_(bf32, a32)
    both = ch.maxabs_many(bfs[0], bf32)
    assert both[0] == many[0] and both[1] == float(np.abs(a32).max())


# ------------------------------------------------------------------------------------------------ device-side Gaussian set!
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,loc", [((40,), (0,)), ((33, 18), (0, 0)), ((33, 18), (1, 0)), ((14, 10, 8), (0, 1, 0))])
def test_set_gaussian_on_the_device(ch, arch, n, loc, dtype):
    """set!(C, grid, (x, y) -> exp(-x^2 - y^2)) (examples/diffusion_2d_mpi.jl:46) evaluated in a kernel at the field's
    coordinates: equals the host evaluation of the same closure (the generic set! path) to 1e-12 relative (CUDA's exp is
    within an ulp of the host's, so not bit for bit); halo and padding stay untouched (field.jl:131-142 writes the interior)."""
    nd = len(n)
    g = ch.UniformGrid(arch, origin=(-1.0,) * nd, extent=(2.0,) * nd, dims=n, dtype=dtype)
    L = tuple(ch.Vertex() if l else ch.Center() for l in loc)
    fd, fh = ch.Field(arch, g, L), ch.Field(arch, g, L)
    ch.set_(fd, g, ch.init_gauss)                                         # device kernel
    ch.set_(fh, g, lambda *x: ch.init_gauss(*x))                          # host evaluation, uploaded
    a, b = fd.parent().astype(np.float64), fh.parent().astype(np.float64)
    tol = 1e-12 if dtype == np.float64 else 3e-7
    assert np.abs(a - b).max() <= tol * np.abs(b).max() and b.max() > 0.3
    inner = tuple(slice(2, -2) for _ in range(nd))
    m = np.ones(a.shape, bool); m[inner] = False
    assert (a[m] == 0).all()


def test_boundary_function_over_mutable_state_is_re_evaluated(ch, arch):
    """The reference calls the boundary function inside every bc! kernel (boundary_function.jl:34-44), so a closure over
    time-dependent state gives the CURRENT value each time; `static=True` declares it pure (values kept)."""
    g = ch.UniformGrid(arch, origin=(0.0, 0.0), extent=(1.0, 1.0), dims=(6, 5))
    f = ch.Field(arch, g, (ch.Vertex(), ch.Center()))
    state = {"t": 1.0}
    for static, want in ((False, 2.0), (True, 1.0)):
        state["t"] = 1.0
        bf = ch.BoundaryFunction(lambda y: state["t"] + 0.0 * y, static=static)
        ch.set_(f, 0.0)
        ch.bc_(arch, g, (f, {"x": ch.Dirichlet(bf)}))
        assert np.all(ch.interior(f)[0, :] == 1.0)
        state["t"] = 2.0
        ch.bc_(arch, g, (f, {"x": ch.Dirichlet(bf)}))
        assert np.all(ch.interior(f)[0, :] == want), (static, ch.interior(f)[0, :])


# ------------------------------------------------------------------------------------------------ large uploads
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,loc,layout", [((200, 180, 160), (0, 1, 0), 0), ((3000, 2100), (1, 0), 0), ((5, 1500, 1400), (0, 0, 1), 0),
                                          ((210, 170, 150), (1, 1, 1), 1)])
def test_large_uploads_take_the_staged_path_and_move_the_same_bytes(ch, arch, n, loc, layout, dtype):
    """set!(f, A) of tens of MB goes through contiguous pieces + a scatter kernel (api.cu, staged upload): interior, a box
    that includes the halo and an off-centre sub-box all read back bit for bit, and nothing outside the box changes."""
    nd = len(n)
    g = ch.UniformGrid(arch, origin=(0.0,) * nd, extent=(1.0,) * nd, dims=n)
    f = ch.Field(arch, g, bloc(ch, loc), dtype=dtype, layout=layout)
    rng = np.random.default_rng(5)
    whole_lo, whole_hi = [-1] * nd, [d + 2 for d in f.dims]
    base = (rng.random(tuple(d + 4 for d in f.dims)) - 0.5).astype(dtype)
    f.from_host(base, whole_lo, whole_hi)                                       # the whole padded array
    assert np.array_equal(f.to_host(whole_lo, whole_hi), base)
    a = ch.pinned_array(arch, f.dims, dtype=dtype) if dtype == np.float64 else np.asfortranarray(rng.random(f.dims).astype(dtype))
    a[...] = rng.random(f.dims)
    ch.set_(f, a)                                                               # the interior
    want = base.copy()
    want[(slice(2, -2),) * nd] = a
    assert np.array_equal(f.to_host(whole_lo, whole_hi), want)
    lo = [3, 2, 5][:nd]
    hi = [d - 4 for d in f.dims]
    if all(h >= l for l, h in zip(lo, hi)):                                      # an off-centre sub-box
        sub = rng.random(tuple(h - l + 1 for l, h in zip(lo, hi))).astype(dtype)
        f.from_host(sub, lo, hi)
        want[tuple(slice(l + 1, h + 2) for l, h in zip(lo, hi))] = sub
        assert np.array_equal(f.to_host(whole_lo, whole_hi), want)
