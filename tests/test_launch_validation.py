"""
CPU-only: the host mirror's flattening of `op => args` against the LIBRARY's own argument rules (chmy_validate_launch:
the checks chmy_launch runs before it touches the device), with descriptor-only fields (Field.shell).  Every solver op in
2D and 3D, every grid operator at every staggered location, both element types, and the refusals: wrong locations, wrong
sizes, aliasing, mixed element types, Float32 fields in the Float64-only solver ops.
"""
import itertools

import numpy as np
import pytest


class _NoArch:
    """UniformGrid only needs the architecture to tell single-device from distributed"""


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    chmy_b200.load_library()
    return chmy_b200


def grid(ch, n, dtype=np.float64):
    nd = len(n)
    return ch.UniformGrid(_NoArch(), origin=(-1.0,) * nd, extent=(2.0,) * nd, dims=n, dtype=dtype)


def F(ch, g, loc=None, dtype=None):
    return ch.Field.shell(g, loc if loc is not None else ch.Center(), dtype)


def vec(ch, g, dtype=None):
    N = g.ndims()
    return ch.FieldTuple(**{"xyz"[D]: F(ch, g, ch.vector_location(D + 1, N), dtype) for D in range(N)})


def ten(ch, g):
    Cn, Vx = ch.Center(), ch.Vertex()
    if g.ndims() == 2:
        return ch.FieldTuple(xx=F(ch, g), yy=F(ch, g), xy=F(ch, g, Vx))
    return ch.FieldTuple(xx=F(ch, g), yy=F(ch, g), zz=F(ch, g), xy=F(ch, g, (Vx, Vx, Cn)), xz=F(ch, g, (Vx, Cn, Vx)),
                         yz=F(ch, g, (Cn, Vx, Vx)))


def loc_of(ch, bits):
    return tuple(ch.Vertex() if b else ch.Center() for b in bits)


def flip(bits, d):
    return tuple(1 - b if a == d else b for a, b in enumerate(bits))


@pytest.mark.parametrize("n", [(12, 10), (12, 10, 8)])
def test_every_solver_op_validates(ch, n):
    g = grid(ch, n)
    nd = len(n)
    L = ch.Launcher(_NoArch(), g, outer_width=(4,) * nd)
    V, rV, qT, tau, tau_old = vec(ch, g), vec(ch, g), vec(ch, g), ten(ch, g), ten(ch, g)
    Pr, dV, T, To = F(ch, g), F(ch, g), F(ch, g), F(ch, g)
    rho_loc = tuple(ch.Vertex() if i == nd - 1 else ch.Center() for i in range(nd))
    rho = F(ch, g, rho_loc)
    par = {k: 0.0 for k in ("x0", "y0", "z0")[:nd]}
    rho_f = ch.FunctionField(ch.init_incl, g, rho_loc, parameters={**par, "r": 0.2, "in": 1.0, "out": 0.0})
    bcV = [(c, {a: (ch.Dirichlet() if a == "xyz"[i] else ch.Neumann()) for a in "xyz"[:nd]}) for i, c in enumerate(V)]
    L.validate(g, (ch.update_old_, (T, tau, To, tau_old)))
    L.validate(g, (ch.update_stress_, (tau, Pr, dV, V, tau_old, 10.0, 0.1, 1.0, 0.07, 0.3, 0.2, g)))
    for r in (rho, rho_f):
        L.validate(g, (ch.update_velocity_, (V, rV, Pr, tau, r, 0.1, 0.01, g)), bc=ch.batch(g, *bcV))
    L.validate(g, (ch.update_thermal_flux_, (qT, T, V, 1e-4, g)))
    L.validate(g, (ch.update_thermal_, (T, To, qT, 0.07, g)), bc=ch.batch(g, (T, ch.Neumann())))
    if nd == 2:
        q, Cf = vec(ch, g), F(ch, g)
        L.validate(g, (ch.compute_q_, (q, Cf, 1.0, g)))
        L.validate(g, (ch.update_C_, (Cf, q, 0.01, g)), bc=ch.batch(g, (Cf, ch.Neumann(2.0))))
    # refusals: a field at the wrong staggered location, a field of another grid's size, a missing rho_g
    with pytest.raises(ch.ChmyError, match="location"):
        L.validate(g, (ch.update_thermal_, (V.x, To, qT, 0.07, g)))
    g2 = grid(ch, tuple(x + 1 for x in n))
    with pytest.raises(ch.ChmyError, match="size"):
        L.validate(g, (ch.update_thermal_, (F(ch, g2), To, qT, 0.07, g)))
    with pytest.raises(ch.ChmyError):
        L.validate(g, (ch.update_velocity_, (V, rV, Pr, tau, None, 0.1, 0.01, g)))
    # the kernels are generic in the element type (test/common.jl:9) -- but one launch is of ONE element type
    g32 = grid(ch, n, np.float32)
    ch.Launcher(_NoArch(), g32).validate(g32, (ch.update_thermal_flux_, (vec(ch, g32), F(ch, g32), vec(ch, g32), 1e-4, g32)))
    with pytest.raises(ch.ChmyError, match="element type"):
        ch.Launcher(_NoArch(), g32).validate(g32, (ch.update_thermal_flux_, (vec(ch, g32), F(ch, g), vec(ch, g32), 1e-4, g32)))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [(9,), (7, 5), (6, 5, 4)])
def test_every_operator_validates_at_every_location(ch, n, dtype):
    g = grid(ch, n, dtype)
    nd = len(n)
    L = ch.Launcher(_NoArch(), g)
    locs = list(itertools.product((0, 1), repeat=nd))
    mk = lambda bits: F(ch, g, loc_of(ch, bits))
    for bits in locs:
        f = mk(bits)
        for dim in range(nd):
            for op in (ch.left_, ch.right_, ch.delta_, ch.partial_):
                L.validate(g, (op(dim + 1), (mk(flip(bits, dim)), f, g)))
                with pytest.raises(ch.ChmyError, match="flipped"):
                    L.validate(g, (op(dim + 1), (mk(bits), f, g)))          # the result lives at the flipped location
            L.validate(g, (ch.partial2_(dim + 1), (mk(bits), f, g)))
            for kb in locs:
                L.validate(g, (ch.dkd_(dim + 1), (mk(bits), f, mk(kb), g)))
        L.validate(g, (ch.lapl_, (mk(bits), f, g)))
        L.validate(g, (ch.divg_grad_, (mk(bits), f, mk(locs[-1]), g)))
        for to in locs:
            L.validate(g, (ch.lerp_, (mk(to), f, g)))
            L.validate(g, (ch.hlerp_, (mk(to), f, g)))
        L.validate(g, (ch.grad_, ([mk(flip(bits, d)) for d in range(nd)], f, g)))
        L.validate(g, (ch.kgrad_, ([mk(flip(bits, d)) for d in range(nd)], f, mk(locs[0]), g)))
        with pytest.raises(ch.ChmyError, match="alias"):
            L.validate(g, (ch.lapl_, (f, f, g)))
    ctr = (0,) * nd
    V = [mk(flip(ctr, d)) for d in range(nd)]
    L.validate(g, (ch.divg_, (mk(ctr), V, g)))
    L.validate(g, (ch.vmag_, (mk(ctr), V, g)))
    if nd > 1:
        with pytest.raises(ch.ChmyError):
            L.validate(g, (ch.divg_, (mk(flip(ctr, 0)), V, g)))
        with pytest.raises(ch.ChmyError, match="Center"):
            L.validate(g, (ch.vmag_, (mk(flip(ctr, 0)), V, g)))
    with pytest.raises(ch.ChmyError, match="dim"):
        L.validate(g, (ch.partial_(3) if nd < 3 else ch.KernelOp("bad", 8, lambda d, s, gg: ([d, s], [], None), 4, 5),
                       (mk(flip(ctr, 0)), mk(ctr), g)))
    other = np.float32 if dtype == np.float64 else np.float64
    with pytest.raises(ch.ChmyError, match="element type"):
        L.validate(g, (ch.lapl_, (mk(ctr), F(ch, g, loc_of(ch, ctr), other), g)))


def test_descriptor_only_fields_are_refused_by_everything_that_touches_storage(ch):
    import ctypes as C
    from chmy_b200 import _lib as L
    g = grid(ch, (8, 6))
    f = F(ch, g)
    info = f.info()
    assert info.origin_ptr is None and info.base_ptr is None and info.dtype == 0 and info.stride[0] == 1 and info.bytes > 0
    ln = C.c_int64()
    L.check(L.lib().chmy_halo_slab_len(f.handle, 0, C.byref(ln)))
    assert ln.value == 6 + 4
    lo, hi = L.i64x3([1, 1]), L.i64x3([8, 6])
    out = C.c_double()
    for rc in (L.lib().chmy_field_fill(None, f.handle, 1.0, lo, hi),
               L.lib().chmy_field_maxabs(None, f.handle, lo, hi, C.byref(out))):
        assert rc != 0                       # NULL context / no storage: an error code, never a crash
    # the batched reduction, through the real ctypes signature: argument checks first, storage check per field
    f2 = F(ch, g)
    hs = (C.c_void_p * 2)(f.handle, f2.handle)
    lo2, hi2 = (C.c_int64 * 6)(1, 1, 0, 1, 1, 0), (C.c_int64 * 6)(8, 6, 0, 8, 6, 0)
    out2 = (C.c_double * 2)()
    fake_ctx = C.c_void_p(8)                 # never dereferenced: every check below fails before the context is used
    lib = L.lib()
    assert lib.chmy_field_maxabs_many(None, 2, hs, lo2, hi2, out2) != 0
    assert lib.chmy_field_maxabs_many(fake_ctx, 0, hs, lo2, hi2, out2) != 0
    assert lib.chmy_field_maxabs_many(fake_ctx, 65, hs, lo2, hi2, out2) != 0
    assert lib.chmy_field_maxabs_many(fake_ctx, 2, hs, lo2, hi2, out2) != 0 and b"descriptor-only" in lib.chmy_last_error()
    hs[1] = None
    assert lib.chmy_field_maxabs_many(fake_ctx, 2, hs, lo2, hi2, out2) != 0
    f2.free()
    f.free()


def test_constant_fields(ch):
    """test/test_fields.jl:56-75, and a ConstantField as the body force of update_velocity! (a chmy_inclusion whose inside
    and outside values coincide)."""
    z, o, v = ch.ZeroField(), ch.OneField(), ch.ValueField(2.0)
    for f, want in ((z, 0.0), (o, 1.0), (v, 2.0)):
        assert f[1, 1, 1] == want and f[2, 2, 2] == want and f.size() == ()
    assert ch.ValueField(0.1, np.float32)[3] == np.float32(0.1)
    for n in ((12, 10), (12, 10, 8)):
        g = grid(ch, n)
        nd = len(n)
        L = ch.Launcher(_NoArch(), g)
        V, rV, tau, Pr = vec(ch, g), vec(ch, g), ten(ch, g), F(ch, g)
        d = L.describe(None, g, (ch.update_velocity_, (V, rV, Pr, tau, ch.ValueField(9.81), 0.1, 0.01, g)))
        assert d.rho_g.active == 1 and d.rho_g.inn == d.rho_g.out == 9.81 and d.rho_g.r == 0.0
        assert [d.rho_g.loc[a] for a in range(nd)] == [1 if a == nd - 1 else 0 for a in range(nd)]
        L.validate(g, (ch.update_velocity_, (V, rV, Pr, tau, ch.ValueField(9.81), 0.1, 0.01, g)))
        L.validate(g, (ch.update_velocity_, (V, rV, Pr, tau, ch.ZeroField(), 0.1, 0.01, g)))


def test_field_sizes_that_cannot_exist_are_refused(ch):
    import ctypes as C
    from chmy_b200 import _lib as L
    h = C.c_void_p()
    big = L.i64x3([1 << 29, 1 << 29, 1 << 29])
    assert L.lib().chmy_field_create_shell(3, big, L.i32x3([0, 0, 0]), 0, 0, C.byref(h)) != 0      # 2^87 elements would wrap around
    assert b"too large" in L.lib().chmy_last_error()
    assert L.lib().chmy_field_create_shell(3, L.i64x3([767, 767, 767]), L.i32x3([0, 0, 0]), 0, 0, C.byref(h)) == 0
    info = L.FieldInfo()
    L.check(L.lib().chmy_field_get_info(h, C.byref(info)))
    assert info.bytes == 8 * (15 + 784 * 771 * 771 + 32) and info.stride[1] == 784 and info.stride[2] == 784 * 771
    L.lib().chmy_field_destroy(h)
    assert L.lib().chmy_field_create_shell(4, big, L.i32x3([0, 0, 0]), 0, 0, C.byref(h)) != 0
    assert L.lib().chmy_field_create_shell(2, L.i64x3([8, 0]), L.i32x3([0, 0]), 0, 0, C.byref(h)) != 0
    assert L.lib().chmy_field_create_shell(2, L.i64x3([8, 8]), L.i32x3([0, 2]), 0, 0, C.byref(h)) != 0
    assert L.lib().chmy_field_create_shell(2, L.i64x3([8, 8]), L.i32x3([0, 1]), 7, 0, C.byref(h)) != 0
    assert L.lib().chmy_field_create_shell(2, L.i64x3([8, 8]), L.i32x3([0, 1]), 0, 9, C.byref(h)) != 0


def test_batches_are_validated_before_the_launch_starts(ch):
    """A descriptor that is wrong in its boundary batches is refused by chmy_validate_launch (= before chmy_launch touches
    the device): field of another grid, Field-valued condition of the wrong dimensionality / element type, mixed element
    types in one batch, an ExchangeBatch on a Bounded side."""
    import ctypes as C
    from chmy_b200 import _lib as L
    g = grid(ch, (12, 10, 8))
    la = ch.Launcher(_NoArch(), g)
    T, To, q = F(ch, g), F(ch, g), vec(ch, g)
    op = (ch.update_thermal_, (T, To, q, 0.1, g))
    la.validate(g, op, bc=ch.batch(g, (T, ch.Neumann())))
    other = F(ch, grid(ch, (13, 10, 8)))
    with pytest.raises(ch.ChmyError, match="does not match the grid"):
        la.validate(g, op, bc=ch.batch(g, (other, ch.Neumann())))
    g2 = grid(ch, (10, 8))                                       # transverse grid of dim x
    la.validate(g, op, bc=ch.batch(g, (T, {"x": ch.Dirichlet(ch.Field.shell(g2, ch.Vertex()))})))
    with pytest.raises(ValueError):                              # a 3D value field for a 3D grid: the mirror refuses it itself
        la.validate(g, op, bc=ch.batch(g, (T, {"x": ch.Dirichlet(F(ch, g))})))
    d = la.describe(None, g, op, bc=ch.batch(g, (T, {"x": ch.Dirichlet(ch.Field.shell(g2, ch.Vertex()))})))
    d.bc[0][0].value_field[0] = T.handle                         # ... and so does the library for a hand-made descriptor
    assert L.lib().chmy_validate_launch(C.byref(d)) != 0 and b"value field must have 2 dims" in L.lib().chmy_last_error()
    with pytest.raises(ch.ChmyError, match="element type"):
        la.validate(g, op, bc=ch.batch(g, (T, {"x": ch.Dirichlet(ch.Field.shell(g2, ch.Vertex(), np.float32))})))
    with pytest.raises(ch.ChmyError, match="too small"):
        la.validate(g, op, bc=ch.batch(g, (T, {"x": ch.Dirichlet(ch.Field.shell(grid(ch, (5, 8)), ch.Vertex()))})))
    T32 = F(ch, grid(ch, (12, 10, 8), np.float32))
    with pytest.raises(ch.ChmyError, match="element type"):
        la.validate(g, op, bc=ch.batch(g, (T, ch.Neumann()), (T32, ch.Neumann())))
    d = la.describe(None, g, op, bc=ch.batch(g, (T, ch.Neumann())))
    d.bc[0][1].kind = L.BATCH_EXCHANGE                           # a hand-made ExchangeBatch on a Bounded side
    assert L.lib().chmy_validate_launch(C.byref(d)) != 0 and b"Bounded" in L.lib().chmy_last_error()
