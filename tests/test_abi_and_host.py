"""CPU-only: the C-ABI library loads and exports every symbol include/chmy_b200.h declares; the host mirror's
pure-host logic (grid numbers, batch normalisation, region algebra, Dims_create) agrees with the oracle's
independent restatement.  No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import chmy_b200
    from chmy_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "chmy_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)                       # prototypes only, not comments
    declared = set(re.findall(r"\b(chmy_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(chmy_b200.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    L.lib()                                   # also verifies ABI version and struct layouts
    assert L.lib().chmy_abi_version() == L.ABI_VERSION == 3


def test_no_fallback_without_gpu():
    """On a box without a GPU the product path must fail loudly, not fall back to anything."""
    import chmy_b200 as ch
    n = C.c_int(0)
    rc = ch.load_library().chmy_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("GPU present")
    with pytest.raises(ch.ChmyError):
        ch.Arch(ch.B200Backend())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chmy.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "chmy_oracle" not in src and "libchmy_oracle" not in src, f


class _FakeArch:
    """host-only stand-in so that UniformGrid (pure numbers) can be built without a device"""


def test_dims_create_matches_oracle(oracle):
    import chmy_b200 as ch
    for nprocs in (1, 2, 3, 4, 6, 8, 12, 16, 24, 64):
        for nd in (1, 2, 3):
            assert ch.dims_create(nprocs, (0,) * nd) == oracle.dims_create(nprocs, (0,) * nd)
    assert ch.dims_create(8, (0, 0, 0)) == (2, 2, 2)
    assert ch.dims_create(4, (0, 0, 0)) == (2, 2, 1)
    assert ch.dims_create(2, (0, 0, 0)) == (2, 1, 1)
    assert ch.dims_create(8, (0, 0)) == (4, 2)
    assert ch.dims_create(8, (0, 4, 0)) == (2, 4, 1)
    with pytest.raises(ch.ChmyError):
        ch.dims_create(7, (2, 0))
    # every process count up to 512: library == oracle, the factors multiply to nprocs and come in non-increasing order
    for nprocs in range(1, 513):
        for nd in (1, 2, 3):
            d = ch.dims_create(nprocs, (0,) * nd)
            assert d == oracle.dims_create(nprocs, (0,) * nd) and int(np.prod(d)) == nprocs
            assert all(a >= b for a, b in zip(d, d[1:])), (nprocs, d)
        for fixed in (2, 3):
            if nprocs % fixed == 0:
                d = ch.dims_create(nprocs, (0, fixed, 0))
                assert d == oracle.dims_create(nprocs, (0, fixed, 0)) and d[1] == fixed and int(np.prod(d)) == nprocs


def test_grid_numbers_match_oracle_and_reference(oracle):
    import chmy_b200 as ch
    g = ch.UniformGrid(_FakeArch(), origin=(-1, -2), extent=(2, 4), dims=(5, 20))       # test_grids.jl:11-19
    og = oracle.Grid((-1, -2), (2, 4), (5, 20))
    assert ch.spacing(g) == og.spacing and ch.inv_spacing(g) == og.inv_spacing
    assert np.allclose(ch.spacing(g), (0.4, 0.2)) and np.allclose(ch.inv_spacing(g), (2.5, 5.0))
    assert g.size(ch.Center()) == (5, 20) and g.size((ch.Center(), ch.Vertex())) == (5, 21)
    for d in range(2):
        for loc, code in ((ch.Center(), 0), (ch.Vertex(), 1)):
            n = g.size(loc)[d]
            for i in (0, 1, 2, n, n + 1):
                assert ch.coord(g, loc, d + 1, i) == og.coord(d, code, i)
    assert np.allclose(ch.bounds(g, ch.Center(), 1), (-0.8, 0.8)) and np.allclose(ch.bounds(g, ch.Center(), 2), (-1.9, 1.9))
    d = g.desc()
    assert d.ndims == 2 and tuple(d.n)[:2] == (5, 20) and d.spacing[1] == og.spacing[1]
    # awkward spacings: the muladd coordinates must agree to the last bit
    g = ch.UniformGrid(_FakeArch(), origin=(-0.3, 0.1, 7.7), extent=(1.7, 2.9, 0.61), dims=(767, 31, 13))
    og = oracle.Grid((-0.3, 0.1, 7.7), (1.7, 2.9, 0.61), (767, 31, 13))
    for d in range(3):
        assert np.array_equal(ch.coords(g, ch.Center(), d + 1), og.coords(d, 0))
        assert np.array_equal(ch.coords(g, ch.Vertex(), d + 1), og.coords(d, 1))


def test_region_algebra_matches_oracle_and_tiles_once(oracle):
    import chmy_b200 as ch
    for n, ow in [((256, 256), (16, 8)), ((767, 767, 767), (128, 8, 4)), ((30, 22, 14), (4, 3, 3))]:
        g = ch.UniformGrid(_FakeArch(), origin=(0,) * len(n), extent=(1,) * len(n), dims=n)
        Lb = ch.Launcher(_FakeArch(), g, outer_width=ow)
        Lo = oracle.Launcher(oracle.Grid((0,) * len(n), (1,) * len(n), n), ow)
        regs = {nm: (lo, hi) for nm, lo, hi in Lo.regions()}
        assert ch.worksize(Lb) == Lo.worksize
        lo = ch.inner_offset(Lb)
        hi = tuple(o + w - 1 for o, w in zip(lo, ch.inner_worksize(Lb)))
        assert regs["inner"] == (tuple(lo), hi)
        total = int(np.prod(ch.inner_worksize(Lb)))
        for D in range(1, len(n) + 1):
            for S in (1, 2):
                off, sz = ch.outer_offset(Lb, D, S), ch.outer_worksize(Lb, D)
                assert regs[f"outer{D-1}{S-1}"] == (off, tuple(o + s - 1 for o, s in zip(off, sz)))
                total += int(np.prod(sz))
        assert total == int(np.prod(ch.worksize(Lb)))                      # the boxes tile the worksize exactly once
    if True:                                                              # explicit coverage count on a small case
        n, ow = (30, 22, 14), (4, 3, 3)
        Lo = oracle.Launcher(oracle.Grid((0,) * 3, (1,) * 3, n), ow)
        cnt = np.zeros(tuple(x + 2 for x in n), dtype=int)
        for _, lo, hi in Lo.regions():
            cnt[tuple(slice(l, h + 1) for l, h in zip(lo, hi))] += 1
        assert (cnt == 1).all()


def test_batch_normalisation_matches_oracle(oracle):
    """batch.jl:72-155: per-axis specs, pruning of `nothing`, Connected sides -> ExchangeBatch/EmptyBatch."""
    import chmy_b200 as ch
    from chmy_b200.boundary_conditions import EmptyBatch, ExchangeBatch, FieldBatch, batch

    class F(ch.Field):                       # a Field without device storage: batch() only shuffles references
        def __init__(self, name):
            self.name, self._h = name, None

    conn = ((ch.Bounded(), ch.Connected()), (ch.Connected(), ch.Connected()), (ch.Bounded(), ch.Bounded()))
    g = ch.StructuredGrid([ch.UniformAxis(0, 1, 8)] * 3, conn)
    og = oracle.Grid((0,) * 3, (1,) * 3, (8,) * 3, [[c.code for c in side] for side in conn])
    vx, vy, vz, T = F("vx"), F("vy"), F("vz"), F("T")

    class OF:
        def __init__(self, name):
            self.name = name
    ox, oy, oz = OF("vx"), OF("vy"), OF("vz")
    oracle_fields = {"vx": ox, "vy": oy, "vz": oz}
    # monkeypatch: the oracle's batch() only needs isinstance(exchange, Field) to be False for tuples
    bs = batch(g, (vx, {"x": ch.Dirichlet(), "y": ch.Neumann(), "z": ch.Neumann()}),
               (vy, {"x": ch.Neumann(), "z": (ch.Dirichlet(2.0), None)}), (vz, {"y": ch.Neumann()}),
               exchange=(vx, vy, vz))
    obs = oracle.batch(og, (ox, {"x": oracle.Dirichlet(), "y": oracle.Neumann(), "z": oracle.Neumann()}),
                       (oy, {"x": oracle.Neumann(), "z": (oracle.Dirichlet(2.0), None)}), (oz, {"y": oracle.Neumann()}),
                       exchange=(ox, oy, oz))
    for D in range(3):
        for S in range(2):
            b, ob = bs[D][S], obs[D][S]
            if ob[0] == "empty":
                assert isinstance(b, EmptyBatch)
            elif ob[0] == "exchange":
                assert isinstance(b, ExchangeBatch) and [f.name for f in b.fields] == [f.name for f in ob[1]]
            else:
                assert isinstance(b, FieldBatch)
                assert [f.name for f in b.fields] == [f.name for f, _ in ob[1]]
                assert [(c.kind, c.value) for c in b.conditions] == [(c.kind, c.value) for _, c in ob[1]]
    assert isinstance(bs[0][0], FieldBatch) and len(bs[0][0].fields) == 2           # x left: vx Dirichlet, vy Neumann
    assert isinstance(bs[0][1], ExchangeBatch) and isinstance(bs[1][0], ExchangeBatch)
    assert isinstance(bs[2][1], FieldBatch) and [f.name for f in bs[2][1].fields] == ["vx"]   # vy right-z pruned
    # no exchange fields on a Connected side -> EmptyBatch; NamedTuple exchange -> per-dim component
    bs2 = batch(g, (T, ch.Neumann()))
    assert isinstance(bs2[1][0], EmptyBatch) and isinstance(bs2[0][0], FieldBatch)
    V = ch.FieldTuple(x=vx, y=vy, z=vz)
    bs3 = batch(g, exchange=V)
    assert [f.name for f in bs3[0][1].fields] == ["vx"] and [f.name for f in bs3[1][1].fields] == ["vy"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_grid_numbers_match_oracle_in_both_element_types(oracle, dtype):
    """UniformAxis{T} on the host mirror rounds exactly where the oracle's (and Julia's) arithmetic does: spacing,
    inv_spacing, vertex/center coordinates (muladd -> one rounding in T), sub-axis arithmetic."""
    import chmy_b200 as ch
    g = ch.UniformGrid(_FakeArch(), origin=(-5.0, -0.3, 1.0), extent=(10.0, 9.1, 0.7), dims=(12, 10, 7), dtype=dtype)
    og = oracle.Grid((-5.0, -0.3, 1.0), (10.0, 9.1, 0.7), (12, 10, 7), dtype=dtype)
    assert g.eltype() == dtype
    for d in range(3):
        ax = g.axes[d]
        assert isinstance(ax.spacing, dtype) and float(ax.spacing) == og.spacing[d] and float(ax.inv_spacing) == og.inv_spacing[d]
        for loc, lc in ((ch.Vertex(), 1), (ch.Center(), 0)):
            for i in range(-1, 16):
                assert float(ax.coord(loc, i)) == og.coord(d, lc, i), (d, lc, i)
            assert float(ax.origin_at(loc)) == og.origin_at(d, lc) and float(ax.extent_at(loc)) == og.extent_at(d, lc)
    desc = g.desc()
    assert desc.spacing[1] == og.spacing[1] and desc.inv_spacing[2] == og.inv_spacing[2]


def test_fma_in_binary32_rounds_once():
    """fma_t(..., float32) is the exact a*b+c rounded once to binary32 (no double rounding through binary64)."""
    from fractions import Fraction
    from chmy_b200.utils import fma_t
    rng = np.random.default_rng(0)
    for _ in range(2000):
        a, b, c = (np.float32(x) for x in (rng.standard_normal(3) * 10.0 ** rng.integers(-3, 4)))
        x = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
        r = fma_t(a, b, c, np.float32)
        lo, hi = np.nextafter(r, np.float32(-np.inf)), np.nextafter(r, np.float32(np.inf))
        assert isinstance(r, np.float32)
        assert abs(Fraction(float(r)) - x) <= min(abs(Fraction(float(lo)) - x), abs(Fraction(float(hi)) - x))
    # a constructed double-rounding trap: exact value just above a binary32 midpoint that binary64 rounds ONTO the midpoint
    a, b = np.float32(1.0 + 2.0 ** -23), np.float32(1.0 + 2.0 ** -23)          # a*b = 1 + 2^-22 + 2^-46
    c = np.float32(2.0 ** -24)                                                   # sum = 1 + 2^-22 + 2^-24 + 2^-46
    x = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    r = fma_t(a, b, c, np.float32)
    assert abs(Fraction(float(r)) - x) < Fraction(1, 2 ** 24)                    # strictly nearest


def test_host_api_surface_matches_the_reference_exports():
    """Every host-side name the reference exports for this path (src/*/…jl `export` lists) exists on the mirror, `!` spelled `_`."""
    import chmy_b200 as ch
    names = ("Architecture SingleDeviceArchitecture Arch get_backend get_device activate_ set_device_ DoubleBuffer swap_ front "
             "back Launcher worksize outer_width inner_worksize inner_offset outer_worksize outer_offset FirstOrderBC Dirichlet "
             "Neumann bc_ BoundaryFunction FieldBatch ExchangeBatch EmptyBatch batch CartesianTopology global_rank shared_rank "
             "node_name cart_comm shared_comm dims cart_coords neighbors neighbor has_neighbor global_size node_size "
             "DistributedArchitecture topology is_gpu_aware exchange_halo_ gather_ AbstractField Field VectorField TensorField "
             "FunctionField location halo interior set_ Location Center Vertex flip Bounded Connected UniformAxis "
             "StructuredGrid UniformGrid nvertices ncenters spacing inv_spacing volume inv_volume coord coords center vertex "
             "centers vertices origin extent bounds axis direction axes_names expand_loc connectivity "
             "left_ right_ delta_ partial_ partial2_ dkd_ lerp_ hlerp_ divg_ divg_grad_ lapl_ vmag_").split()
    missing = [n for n in names if not hasattr(ch, n)]
    assert not missing, missing
    g = ch.UniformGrid(_FakeArch(), origin=(0.0, 0.0), extent=(1.0, 2.0), dims=(4, 5))
    assert (ch.nvertices(g, 1), ch.ncenters(g, 2)) == (5, 5) and ch.axis(g, 2) is g.axes[1]
    assert ch.vertex(g, 1, 2) == 0.25 and ch.center(g, 2, 1) == 0.2 and ch.direction(g, "y") == 2
    assert ch.volume(g, ch.Center(), 1, 1) == 0.25 * 0.4 and ch.inv_volume(g, ch.Vertex(), 1, 1) == 4.0 * 2.5
    db = ch.DoubleBuffer("a", "b")
    ch.swap_(db)
    assert (ch.front(db), ch.back(db)) == ("b", "a")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_function_field_host_values(oracle, dtype):
    """FunctionField with an arbitrary body (function_field.jl:49-59): the host evaluation that stands in for the in-kernel
    closure covers every index a kernel can read (0..d+1) at the oracle's coordinates; discrete bodies get (grid, loc, I...)."""
    import chmy_b200 as ch
    g = ch.UniformGrid(_FakeArch(), origin=(-1.0, 0.5), extent=(2.0, 1.5), dims=(6, 5), dtype=dtype)
    og = oracle.Grid((-1.0, 0.5), (2.0, 1.5), (6, 5), dtype=dtype)
    loc = (ch.Center(), ch.Vertex())
    ff = ch.FunctionField(lambda x, y, a: a * x + y * y, g, loc, parameters=(3.0,))
    assert not ff.in_kernel() and ff.dims == (6, 6)
    v = ff.values()
    assert v.shape == (8, 8) and v.dtype == dtype and v.flags.f_contiguous
    for i in (0, 1, 3, 7):
        for j in (0, 2, 7):
            x, y = dtype(og.coord(0, 0, i)), dtype(og.coord(1, 1, j))
            assert v[i, j] == dtype(3.0) * x + y * y
    dff = ch.FunctionField(lambda grid, l, i, j, s: s * (10 * i + j), g, loc, discrete=True, parameters=(0.5,))
    w = dff.values()
    assert w[3, 4] == 0.5 * 34 and w[0, 7] == 0.5 * 7 and w.dtype == dtype
    inc = ch.FunctionField(ch.init_incl, g, loc, parameters={"x0": 0.0, "y0": 1.0, "r": 0.3, "in": 1.0, "out": 0.0})
    assert inc.in_kernel() and inc.inclusion().r == 0.3
