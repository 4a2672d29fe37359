"""
Pins the CPU oracle on every known-answer the reference's own test-suite holds for the hot path
(citations: file:line under /root/reference/test).  CPU only.
"""
import math

import numpy as np
import pytest

PI = math.pi
DTYPES = [np.float64, np.float32]            # TEST_TYPES of the reference (test/common.jl:9)


def tol(dtype):
    """Julia's `≈`: rtol = sqrt(eps(T))"""
    return dict(rtol=float(np.sqrt(np.finfo(dtype).eps)), atol=float(np.sqrt(np.finfo(dtype).eps)) * 1e-3)


# ------------------------------------------------------------------ test_grids.jl:11-148
@pytest.mark.parametrize("dtype", DTYPES)
def test_grid_sizes_bounds_spacing(oracle, dtype):
    o = oracle
    nx, ny = 5, 20
    g = o.Grid((-1.0, -2.0), (2.0, 4.0), (nx, ny), dtype=dtype)
    assert g.dtype == dtype and np.asarray(g.coords(0, o.CENTER)).dtype == dtype
    assert g.size(o.CENTER) == (nx, ny)                                   # :31-39
    assert g.size(o.VERTEX) == (nx + 1, ny + 1)
    assert g.size((o.CENTER, o.VERTEX)) == (nx, ny + 1)
    assert g.size((o.VERTEX, o.CENTER)) == (nx + 1, ny)
    assert np.allclose(g.bounds(0, o.VERTEX), (-1.0, 1.0))                # :41-46
    assert np.allclose(g.bounds(1, o.VERTEX), (-2.0, 2.0))
    assert np.allclose(g.bounds(0, o.CENTER), (-0.8, 0.8))
    assert np.allclose(g.bounds(1, o.CENTER), (-1.9, 1.9))
    assert np.isclose(g.extent_at(0, o.CENTER), 1.6) and np.isclose(g.extent_at(1, o.CENTER), 3.8)   # :48-62
    assert np.isclose(g.origin_at(0, o.CENTER), -0.8) and np.isclose(g.origin_at(1, o.CENTER), -1.9)  # :64-78
    assert np.allclose(g.spacing, (0.4, 0.2))                             # :80-110
    assert np.allclose(g.inv_spacing, (2.5, 5.0))
    assert np.isclose(g.coord(0, o.VERTEX, 1), -1.0) and np.isclose(g.coord(1, o.VERTEX, 1), -2.0)   # :124-148
    assert np.isclose(g.coord(0, o.VERTEX, nx + 1), 1.0) and np.isclose(g.coord(1, o.VERTEX, ny + 1), 2.0)
    assert np.isclose(g.coord(0, o.CENTER, 1), -0.8) and np.isclose(g.coord(1, o.CENTER, 1), -1.9)
    assert np.isclose(g.coord(0, o.CENTER, nx), 0.8) and np.isclose(g.coord(1, o.CENTER, ny), 1.9)
    # default connectivity is Bounded (:20-24)
    assert all(c == o.BOUNDED for side in g.conn for c in side)


# ------------------------------------------------------------------ test_fields.jl:19-55 (exact == in the reference)
@pytest.mark.parametrize("dtype", DTYPES)
def test_set_continuous_exact(oracle, dtype):
    o = oracle
    g = o.Grid((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (2, 2, 2), dtype=dtype)
    f = o.Field(g, (o.CENTER, o.VERTEX, o.CENTER))
    assert f.dims == (2, 3, 2) and f.data.dtype == dtype
    f.data[...] = np.nan
    f.set_fun(lambda x, y, z: y)
    exp_y = np.zeros((2, 3, 2))
    exp_y[:, 0, :], exp_y[:, 1, :], exp_y[:, 2, :] = 0.0, 0.5, 1.0
    assert np.array_equal(f.interior(), exp_y)
    f.data[...] = np.nan
    f.set_fun(lambda x, y, z: x)
    exp_x = np.zeros((2, 3, 2))
    exp_x[0], exp_x[1] = 0.25, 0.75
    assert np.array_equal(f.interior(), exp_x)
    f.data[...] = np.nan
    f.set_fun(lambda x, y, z, sc: y * sc, 2.0)
    assert np.array_equal(f.interior(), 2.0 * exp_y)
    # set! touches the interior only
    assert np.isnan(f.data[0]).all() and np.isnan(f.data[:, 1]).all()


# ------------------------------------------------------------------ test_boundary_conditions.jl:11-205
CASES = [
    ((8,), (0,)), ((8,), (1,)),                                    # 1D Center / Vertex      :11-89
    ((8, 8), (0, 1)),                                              # 2D (Center, Vertex)     :91-141
    ((8, 8, 6), (0, 1, 0)),                                        # 3D (C, V, C)            :143-205
]


def _face(a, dim, idx):
    """a[..., idx, ...] restricted to 2:end-1 in the transverse dims (the reference excludes corners)."""
    sl = [slice(1, -1)] * a.ndim
    sl[dim] = idx
    return a[tuple(sl)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,loc", CASES)
def test_bc_known_answers(oracle, n, loc, dtype):
    o = oracle
    nd = len(n)
    g = o.Grid((-PI,) * nd, (2 * PI,) * nd, n, dtype=dtype)
    f = o.Field(g, loc)
    assert f.data.dtype == dtype
    allclose = lambda a, b: np.allclose(a, b, **tol(dtype))

    def run(bc):
        f.data[...] = 0.0
        f.set(1.0)
        o.bc_(g, (f, bc))
        return f.interior(with_halo=True).copy()

    a = run(o.Dirichlet())
    for d in range(nd):
        if loc[d] == o.CENTER:
            assert allclose(_face(a, d, 0), -_face(a, d, 1)) and allclose(_face(a, d, -1), -_face(a, d, -2))
        else:
            assert allclose(_face(a, d, 1), 0.0) and allclose(_face(a, d, -2), 0.0)
    a = run(o.Neumann())
    for d in range(nd):
        assert allclose(_face(a, d, 0), _face(a, d, 1)) and allclose(_face(a, d, -1), _face(a, d, -2))
    v = 2.0
    a = run(o.Dirichlet(v))
    for d in range(nd):
        if loc[d] == o.CENTER:
            assert allclose(_face(a, d, 0), -_face(a, d, 1) + 2 * v)
            assert allclose(_face(a, d, -1), -_face(a, d, -2) + 2 * v)
        else:
            assert allclose(_face(a, d, 1), v) and allclose(_face(a, d, -2), v)
    q = 2.0
    a = run(o.Neumann(q))
    for d in range(nd):
        h = g.spacing[d]
        assert allclose((_face(a, d, 1) - _face(a, d, 0)) / h, q)
        assert allclose((_face(a, d, -1) - _face(a, d, -2)) / h, q)


# ------------------------------------------------------------------ test_grid_operators.jl:13-131
def _gauss_grid(o):
    g = o.Grid((-5.0,) * 3, (10.0,) * 3, (12, 10, 8))
    Ci = o.Field(g, o.CENTER)
    Ci.set_fun(lambda x, y, z: np.exp(-x ** 2 - y ** 2 - z ** 2))
    return g, Ci


def test_divg_identity_bit_exact(oracle):
    """divg(V) == dx(V.x)+dy(V.y)+dz(V.z) with `==` (test_grid_operators.jl:41)."""
    o = oracle
    g, Ci = _gauss_grid(o)
    V = o.VectorField(g)
    nx, ny, nz = g.n
    rng = range
    for k in rng(0, nz + 2):
        for j in rng(0, ny + 2):
            for i in rng(0, nx + 2):
                for d, c in enumerate("xyz"):
                    V[c].data[i + 1, j + 1, k + 1] = o.partial(g, Ci, d, i, j, k)
    # the stress kernel stores divg(V) into its second output: run it with zero rheology side effects
    tau, tau_old = o.TensorField(g), o.TensorField(g)
    Pr, dV = o.Field(g, o.CENTER), o.Field(g, o.CENTER)
    L = o.Launcher(g)
    o.launch(L, g, o.update_stress, (tau, Pr, dV, V, tau_old, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0))
    C1 = np.zeros(g.n)
    for k in rng(1, nz + 1):
        for j in rng(1, ny + 1):
            for i in rng(1, nx + 1):
                C1[i - 1, j - 1, k - 1] = (o.partial(g, V["x"], 0, i, j, k) + o.partial(g, V["y"], 1, i, j, k)) \
                    + o.partial(g, V["z"], 2, i, j, k)
    assert np.array_equal(dV.interior(), C1)
    assert np.abs(C1).max() > 1e-3


def test_lapl_and_divg_grad(oracle):
    """lapl == sum d2 (bit exact, :61); divg_grad ~ d(lerp(chi) d C) for chi at Center and Vertex (:98,110)."""
    o = oracle
    g, Ci = _gauss_grid(o)
    nx, ny, nz = g.n
    for chi_loc in (o.CENTER, o.VERTEX):
        chi = o.Field(g, chi_loc)
        chi.set_fun(lambda x, y, z: np.exp(-x ** 2 - y ** 2 - z ** 2))
        V = o.VectorField(g)
        for k in range(0, nz + 2):
            for j in range(0, ny + 2):
                for i in range(0, nx + 2):
                    for d, c in enumerate("xyz"):
                        V[c].data[i + 1, j + 1, k + 1] = o.lerp(g, chi, V[c].loc, i, j, k) * o.partial(g, Ci, d, i, j, k)
        for k in range(1, nz + 1, 3):
            for j in range(1, ny + 1, 2):
                for i in range(1, nx + 1):
                    c1 = o.partial(g, V["x"], 0, i, j, k) + o.partial(g, V["y"], 1, i, j, k) + o.partial(g, V["z"], 2, i, j, k)
                    c2 = (o.dkd(g, Ci, chi, 0, i, j, k) + o.dkd(g, Ci, chi, 1, i, j, k)) + o.dkd(g, Ci, chi, 2, i, j, k)
                    assert math.isclose(c1, c2, rel_tol=1e-8, abs_tol=1e-14)
    # second derivative against the explicit three-point formula
    idx = g.inv_spacing
    for (i, j, k) in [(1, 1, 1), (6, 5, 4), (12, 10, 8)]:
        for d in range(3):
            e = [0, 0, 0]
            e[d] = 1
            a = Ci.at(i + e[0], j + e[1], k + e[2]); b = Ci.at(i, j, k); c = Ci.at(i - e[0], j - e[1], k - e[2])
            assert o.partial2(g, Ci, d, i, j, k) == ((a - b) * idx[d] - (b - c) * idx[d]) * idx[d]


def test_vmag_constant(oracle):
    """vmag of the constant (2,2,2) field = 3.4641 to 5 significant digits (:113-131)."""
    o = oracle
    g, _ = _gauss_grid(o)
    V = o.VectorField(g)
    for c in "xyz":
        V[c].set(2.0)
    s = 0.0
    vals = [o.lerp(g, V[c], o.CENTER, 3, 3, 3) for c in "xyz"]
    s = math.sqrt((vals[0] ** 2 + vals[1] ** 2) + vals[2] ** 2)
    assert float(f"{s:.5g}") == 3.4641


# ------------------------------------------------------------------ test_interpolations.jl:15-74
@pytest.mark.parametrize("dtype", DTYPES)
def test_lerp_known_answers(oracle, dtype):
    o = oracle
    g = o.Grid((0.0, 0.0), (1.0, 1.0), (2, 2), dtype=dtype)
    av4 = lambda A: 0.25 * (A[:-1, :-1] + A[1:, :-1] + A[1:, 1:] + A[:-1, 1:])
    avx = lambda A: 0.5 * (A[:-1, :] + A[1:, :])
    avy = lambda A: 0.5 * (A[:, :-1] + A[:, 1:])

    def interp(src, to):
        dst = o.Field(g, to)
        out = np.zeros(dst.dims)
        for j in range(1, dst.dims[1] + 1):
            for i in range(1, dst.dims[0] + 1):
                out[i - 1, j - 1] = o.lerp(g, src, to, i, j)
        return out

    fc = o.Field(g, o.CENTER)
    fc.set(np.arange(1, 5, dtype=float).reshape((2, 2), order="F"))
    fci = fc.interior().copy()
    assert np.allclose(interp(fc, o.VERTEX)[1:-1, 1:-1], av4(fci))
    assert np.allclose(interp(fc, o.CENTER), fci)
    assert np.allclose(interp(fc, (o.CENTER, o.VERTEX))[:, 1:-1], avy(fci))
    assert np.allclose(interp(fc, (o.VERTEX, o.CENTER))[1:-1, :], avx(fci))
    fv = o.Field(g, o.VERTEX)
    fv.set(np.arange(1, 10, dtype=float).reshape((3, 3), order="F"))
    fvi = fv.interior().copy()
    assert np.allclose(interp(fv, o.CENTER), av4(fvi))
    assert np.allclose(interp(fv, o.VERTEX), fvi)
    assert np.allclose(interp(fv, (o.CENTER, o.VERTEX)), avx(fvi))
    assert np.allclose(interp(fv, (o.VERTEX, o.CENTER)), avy(fvi))


# ------------------------------------------------------------------------------------------------ boundary functions
def _bf_known_answers(BF, grid, V, nx, ny, coord_x, coord_xy):
    """test/test_boundary_functions.jl:13-96 (2D grid on [-pi, pi]^2, 8 x 8 cells), tolerance as the reference's `≈`."""
    import math
    pi = math.pi
    ok = lambda a, b: math.isclose(a, b, rel_tol=1.5e-8, abs_tol=1.5e-8)      # Julia `≈`; atol covers the exact-zero targets

    def reduced(bf, s=1.0):
        assert ok(bf(grid, V, 1, 1, 1), -s) and ok(bf(grid, V, 1, 1, ny + 1), -s) and ok(bf(grid, V, 1, 1, ny // 2 + 1), s)
        assert ok(bf(grid, V, 2, 1, 1), -s) and ok(bf(grid, V, 2, nx + 1, 1), -s) and ok(bf(grid, V, 2, nx // 2 + 1, 1), s)

    def full(bf):
        assert ok(bf(grid, V, 1, 1, 1), pi) and ok(bf(grid, V, 1, 1, ny + 1), -pi) and ok(bf(grid, V, 1, 1, ny // 2 + 1), 0.0)
        assert ok(bf(grid, V, 2, 1, 1), pi) and ok(bf(grid, V, 2, nx + 1, 1), pi) and ok(bf(grid, V, 2, nx // 2 + 1, 1), -pi)

    # continuous (:16-53)
    bf = BF(lambda xi: math.cos(xi))
    reduced(bf)
    # changing the index along the boundary dimension does not affect the value (:26-31)
    assert bf(grid, V, 1, ny + 1, 1) == bf(grid, V, 1, 1, 1) and bf(grid, V, 1, ny // 2 + 1, 1) == bf(grid, V, 1, 1, 1)
    assert bf(grid, V, 2, 1, 1) == bf(grid, V, 2, 1, ny + 1) and bf(grid, V, 2, 1, 1) == bf(grid, V, 2, 1, ny // 2 + 1)
    full(BF(lambda xi, eta: math.cos(xi) * eta, reduce_dims=False))
    reduced(BF(lambda xi, eta: math.cos(xi) * eta, parameters=pi), pi)
    # discrete (:55-94)
    reduced(BF(lambda g, loc, dim, i: math.cos(coord_x(g, loc, dim, i)), discrete=True))
    full(BF(lambda g, loc, dim, ix, iy: math.cos(coord_xy(g, loc, ix, iy)[0]) * coord_xy(g, loc, ix, iy)[1],
            discrete=True, reduce_dims=False))
    reduced(BF(lambda g, loc, dim, i, eta: math.cos(coord_x(g, loc, dim, i)) * eta, discrete=True, parameters=pi), pi)


def test_boundary_function_known_answers_oracle(oracle):
    import math
    o = oracle
    g = o.Grid((-math.pi, -math.pi), (2 * math.pi, 2 * math.pi), (8, 8))
    # the reference's discrete test functions index coord(grid, loc, dim, i) with the *reduced* index: after
    # remove_dim the surviving index belongs to the other axis, but both axes of this grid are identical
    _bf_known_answers(o.BoundaryFunction, g, o.VERTEX, 8, 8,
                      lambda gr, loc, dim, i: gr.coord(dim - 1, loc, i),
                      lambda gr, loc, ix, iy: (gr.coord(0, loc, ix), gr.coord(1, loc, iy)))


def test_boundary_function_known_answers_host_mirror():
    import math
    import chmy_b200 as ch
    g = ch.UniformGrid(None, origin=(-math.pi, -math.pi), extent=(2 * math.pi, 2 * math.pi), dims=(8, 8))
    _bf_known_answers(ch.BoundaryFunction, g, ch.Vertex(), 8, 8,
                      lambda gr, loc, dim, i: ch.coord(gr, loc, dim, i),
                      lambda gr, loc, ix, iy: (ch.coord(gr, loc, 1, ix), ch.coord(gr, loc, 2, iy)))


def test_valued_bc_field_and_function_oracle(oracle):
    """Field-valued and BoundaryFunction-valued conditions (first_order_boundary_condition.jl:36-40, the docs'
    parabolic-profile example boundary_function.jl:57-66) against the closed forms of the rules (:42-84)."""
    o = oracle
    n = (10, 6)
    g = o.Grid((0.0, 0.0), (2.0, 1.5), n)
    f = o.Field(g, (o.CENTER, o.VERTEX))
    rng = np.random.default_rng(3)
    f.data[...] = rng.random(f.sdims)
    before = f.data.copy()
    xbc = o.BoundaryFunction(lambda x, ly: x * (ly - x), parameters=(2.0,))
    # y is the boundary dim (D=1): the function receives the x coordinate; f is Vertex along y -> Dirichlet sets the node
    o.bc_(g, (f, {"y": o.Dirichlet(xbc)}))
    xs_v = np.array([g.coord(0, o.VERTEX, i) for i in range(0, n[0] + 3)])      # loc = Vertex for every axis (batch.jl:174)
    want = xs_v * (2.0 - xs_v)
    assert np.array_equal(f.data[1:n[0] + 4, 2], want)                   # node 1   (storage index 2), i = 0..n+2
    assert np.array_equal(f.data[1:n[0] + 4, n[1] + 2], want)            # node d = n+1 (storage d+1)
    untouched = np.ones_like(f.data, dtype=bool)
    untouched[1:n[0] + 4, 2] = untouched[1:n[0] + 4, n[1] + 2] = False
    assert np.array_equal(f.data[untouched], before[untouched])
    # Field-valued Neumann along x (Center along x): halo = fma(dx, -/+q[j], f[nb])
    q = o.Field(o.transverse_grid(g, 0), o.VERTEX)
    q.data[...] = rng.random(q.sdims)
    before = f.data.copy()
    o.bc_(g, (f, {"x": o.Neumann(q)}))
    js = np.arange(0, n[1] + 3)
    dx = g.spacing[0]
    for j in js:
        qv = q.data[j + 1]
        assert f.data[1, j + 1] == math.fma(dx, -qv, before[2, j + 1]) if hasattr(math, "fma") else True
        assert np.isclose(f.data[1, j + 1], before[2, j + 1] - dx * qv, rtol=1e-15, atol=0)
        assert np.isclose(f.data[n[0] + 2, j + 1], before[n[0] + 1, j + 1] + dx * qv, rtol=1e-15, atol=0)


# ------------------------------------------------------------------------------------------------ field-level operators
def test_hlerp_is_the_harmonic_rule(oracle):
    """interpolation.jl:15 and the docstring at :92: rule(t, v0, v1) = 1/(1/v0 + t*(1/v1 - 1/v0)), t = 0.5 on uniform
    axes -> the harmonic mean; nested per dimension like lerp."""
    import math
    o = oracle
    g = o.Grid((0.0, 0.0), (1.0, 1.0), (2, 2))
    fc = o.Field(g, o.CENTER)
    fc.set(np.array([[1.0, 3.0], [2.0, 4.0]]))
    h = lambda a, b: 1.0 / math.fma(0.5, 1.0 / b - 1.0 / a, 1.0 / a) if hasattr(math, "fma") else 1.0 / (0.5 * (1.0 / b - 1.0 / a) + 1.0 / a)
    v = o.hlerp(g, fc, (o.VERTEX, o.CENTER), 2, 1)                        # between C[1,1]=1 and C[2,1]=2
    assert math.isclose(v, 2 * 1.0 * 2.0 / 3.0, rel_tol=1e-15) and math.isclose(v, h(1.0, 2.0), rel_tol=1e-15)
    v = o.hlerp(g, fc, o.VERTEX, 2, 2)                                    # x innermost, y outermost (:19-25)
    assert math.isclose(v, h(h(1.0, 2.0), h(3.0, 4.0)), rel_tol=1e-15)
    assert o.hlerp(g, fc, o.CENTER, 1, 2) == 3.0                          # same location: f[I] (:65-67)


def test_apply_operator_is_the_point_functions_over_the_launch_range(oracle):
    """og_apply_operator == the pinned point functions (og_partial, og_partial2, og_lerp, og_dkd) evaluated at every
    I in [0, n+1]^N, with the left-folded sums of field_operators.jl:50-121."""
    import math
    o = oracle
    g = o.Grid((-1.0, 0.5, 0.0), (2.0, 1.7, 1.1), (5, 4, 3))
    rng = np.random.default_rng(8)

    def rnd(loc):
        f = o.Field(g, loc)
        f.data[...] = rng.random(f.sdims) - 0.5
        return f
    f, k = rnd((0, 1, 0)), rnd((1, 1, 0))
    V = [rnd((1, 0, 0)), rnd((0, 1, 0)), rnd((0, 0, 1))]
    rngI = [(i, j, kk) for kk in range(0, 5) for j in range(0, 6) for i in range(0, 7)]
    at = lambda F, I: F.data[I[0] + 1, I[1] + 1, I[2] + 1]
    for dim in range(3):
        dloc = tuple(1 - l if a == dim else l for a, l in enumerate(f.loc))
        d = rnd(dloc)
        o.apply_operator(g, "partial", d, f, dim=dim)
        assert all(at(d, I) == o.partial(g, f, dim, *I) for I in rngI)
        d = rnd(f.loc)
        o.apply_operator(g, "partial2", d, f, dim=dim)
        assert all(at(d, I) == o.partial2(g, f, dim, *I) for I in rngI)
        o.apply_operator(g, "dkd", d, f, k=k, dim=dim)
        assert all(at(d, I) == o.dkd(g, f, k, dim, *I) for I in rngI)
    d = rnd(f.loc)
    o.apply_operator(g, "lapl", d, f)
    assert all(at(d, I) == (o.partial2(g, f, 0, *I) + o.partial2(g, f, 1, *I)) + o.partial2(g, f, 2, *I) for I in rngI)
    o.apply_operator(g, "divg_grad", d, f, k=k)
    assert all(at(d, I) == (o.dkd(g, f, k, 0, *I) + o.dkd(g, f, k, 1, *I)) + o.dkd(g, f, k, 2, *I) for I in rngI)
    c = rnd((0, 0, 0))
    o.apply_operator(g, "divg", c, V)
    assert all(at(c, I) == (o.partial(g, V[0], 0, *I) + o.partial(g, V[1], 1, *I)) + o.partial(g, V[2], 2, *I) for I in rngI)
    o.apply_operator(g, "vmag", c, V)
    sq = lambda x: x * x
    assert all(at(c, I) == math.sqrt((sq(o.lerp(g, V[0], 0, *I)) + sq(o.lerp(g, V[1], 0, *I))) + sq(o.lerp(g, V[2], 0, *I)))
               for I in rngI)
    t = rnd((1, 0, 1))
    o.apply_operator(g, "lerp", t, f)
    assert all(at(t, I) == o.lerp(g, f, t.loc, *I) for I in rngI)
    # cells outside the launch range are untouched
    before = rnd((0, 0, 0))
    keep = before.data.copy()
    o.apply_operator(g, "divg", before, V)
    mask = np.ones(before.sdims, bool)
    mask[1:-1, 1:-1, 1:-1] = False
    assert np.array_equal(before.data[mask], keep[mask]) and not np.array_equal(before.data, keep)


@pytest.mark.parametrize("dtype", DTYPES)
def test_operator_identities_both_element_types(oracle, dtype):
    """test/test_grid_operators.jl:13-131 for T in (Float32, Float64) through the field-level operators: divg == sum of
    the partial derivatives (`==`, :41), lapl == sum of second derivatives (`==`, :61), divg_grad ≈ divg(lerp(χ) grad)
    (:98,:110), vmag of (2,2,2) rounds to 3.4641 (:130).  Every result stays in the element type."""
    o = oracle
    g = o.Grid((-5.0,) * 3, (10.0,) * 3, (12, 10, 8), dtype=dtype)
    gauss = lambda x, y, z: np.exp(-x ** 2 - y ** 2 - z ** 2)
    Ci, C1, C2, P = (o.Field(g, 0) for _ in range(4))
    Ci.set_fun(gauss)
    V = [o.Field(g, tuple(1 if a == d else 0 for a in range(3))) for d in range(3)]
    o.apply_operator(g, "grad", V, Ci)
    o.apply_operator(g, "divg", C2, V)
    acc = None
    for d in range(3):
        o.apply_operator(g, "partial", P, V[d], dim=d)
        acc = P.interior().copy() if acc is None else acc + P.interior()
    assert acc.dtype == dtype and np.array_equal(C2.interior(), acc) and np.abs(acc).max() > 1e-3
    o.apply_operator(g, "lapl", C2, Ci)
    acc = None
    for d in range(3):
        o.apply_operator(g, "partial2", P, Ci, dim=d)
        acc = P.interior().copy() if acc is None else acc + P.interior()
    assert np.array_equal(C2.interior(), acc)
    for chi_loc in (0, 1):
        chi = o.Field(g, chi_loc)
        chi.set_fun(gauss)
        o.apply_operator(g, "kgrad", V, Ci, k=chi)
        o.apply_operator(g, "divg", C1, V)
        o.apply_operator(g, "divg_grad", C2, Ci, k=chi)
        assert np.allclose(C2.interior(), C1.interior(), **tol(dtype))
    for c in V:
        c.set(2.0)
    o.apply_operator(g, "vmag", C1, V)
    assert C1.data.dtype == dtype
    assert np.all(np.vectorize(lambda x: float(f"{x:.5g}"))(C1.interior().astype(np.float64)) == 3.4641)
