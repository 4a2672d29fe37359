"""
CPU proof of the boundary-condition and halo-slab kernels (chmy.jl_b200/csrc/bc_kernels.cuh -> k_bc_dim, k_slab of bc.cu).

The per-point bodies are plain C++ shared by nvcc and the host compiler; tests/emul/bc_emul.cpp calls them once per thread
of the grid bc.cu launches, on host arrays in the library's PITCHED layout.  The descriptors (BcBatchDev, SlabBatch) are
filled here the way run_bc_dim / make_slab_batch fill them.  Every rule of
src/BoundaryConditions/first_order_boundary_condition.jl:34-84 (Dirichlet on Vertex / Center, Neumann; value nothing |
Number | lower-dimensional Field | BoundaryFunction), on every side, dimension and staggered location in 1D / 2D / 3D, and
every send / receive slab of src/Distributed/communication_views.jl:1-34 must be BIT-identical to the oracle
(oracle/chmy_oracle.c: og_bc_apply, og_bc_apply_field, og_pack_send, og_unpack_recv -- pinned on the reference's own tests
by tests/test_oracle_golden.py and tests/test_oracle_transliteration.py), in both element types (test/common.jl:9).
The -m gpu suite then checks the compiled kernels against the same oracle.
"""
import ctypes as C
import itertools

import numpy as np
import pytest

from test_operators_emulation import Pitched

MAXF = 8          # BCK_MAX_FIELDS == CHMY_MAX_BATCH_FIELDS
ALLF = 6          # BCK_ALL_FIELDS
DTYPES = [np.float64, np.float32]


class BckView(C.Structure):
    _fields_ = [("p", C.c_void_p), ("sy", C.c_longlong), ("sz", C.c_longlong)]


def _structs(R):
    class BcEntry(C.Structure):
        _fields_ = [("f", BckView), ("kind", C.c_int), ("vertex", C.c_int), ("d", C.c_int), ("side", C.c_int),
                    ("value", R), ("vp", C.c_void_p), ("vsy", C.c_longlong)]

    class BcBatchDev(C.Structure):
        _fields_ = [("n", C.c_int), ("dim", C.c_int), ("nt", C.c_int * 2), ("spacing", R), ("e", BcEntry * (2 * MAXF))]

    class BcRule(C.Structure):
        _fields_ = [("kind", C.c_int), ("value", R), ("vp", C.c_void_p), ("vsy", C.c_longlong)]

    class BcAllField(C.Structure):
        _fields_ = [("f", BckView), ("d", C.c_int * 3), ("vertex", C.c_int * 3), ("r", (BcRule * 2) * 3)]

    class BcAllDev(C.Structure):
        _fields_ = [("nf", C.c_int), ("nd", C.c_int), ("n", C.c_int * 3), ("spacing", R * 3), ("fld", BcAllField * ALLF),
                    ("nact", C.c_int), ("act", C.c_ubyte * (6 * ALLF))]

    class SlabEntry(C.Structure):
        _fields_ = [("f", BckView), ("idx", C.c_int), ("e0", C.c_int), ("e1", C.c_int), ("off", C.c_longlong)]

    class SlabBatch(C.Structure):
        _fields_ = [("n", C.c_int), ("dim", C.c_int), ("nd", C.c_int), ("e", SlabEntry * MAXF)]

    return BcBatchDev, SlabBatch, BcAllDev


S64 = _structs(C.c_double)
S32 = _structs(C.c_float)


@pytest.fixture(scope="module")
def emul():
    from helpers import build_emul
    lib = build_emul("bc_emul")
    for f32, (B, S, A) in enumerate((S64, S32)):
        assert lib.bc_emul_sizeof(0, f32) == C.sizeof(B) and lib.bc_emul_sizeof(1, f32) == C.sizeof(S)
        assert lib.bc_emul_sizeof(2, f32) == C.sizeof(A)
    return lib


def view_of(p: Pitched) -> BckView:
    o = p.opr()
    return BckView(o.p, o.sy, o.sz)


def flip(loc, d):
    return tuple(1 - l if a == d else l for a, l in enumerate(loc))


def rnd_field(o, g, loc, rng):
    f = o.Field(g, loc)
    f.data[...] = rng.random(f.sdims) - 0.5               # interior, halo AND padding
    return f


def same_bits(a, b):
    return ((a == b) & (np.signbit(a) == np.signbit(b))) | (np.isnan(a) & np.isnan(b))


# ---------------------------------------------------------------------------------------------- boundary conditions

def run_bc_both(o, emul, g, D, sides):
    """sides = (left, right), each a list of (oracle field, BC) or None.  Applies dimension D through the oracle (side 1
    then side 2, fields in batch order: batch.jl:20-29,163-184) and through ONE emulated k_bc_dim launch; compares bits of
    the whole padded arrays."""
    f32 = g.dtype == np.float32
    B = (S32 if f32 else S64)[0]
    b = B()
    b.dim, b.spacing = D, g.spacing[D]
    keep = []                                             # Pitched copies (and value fields) stay alive over the call
    pit = {}
    for s in range(2):
        for f, bc in sides[s] or []:
            if id(f) not in pit:
                pit[id(f)] = (f, Pitched(f))
    for s in range(2):
        for f, bc in sides[s] or []:
            e = b.e[b.n]
            b.n += 1
            e.f = view_of(pit[id(f)][1])
            e.kind, e.vertex, e.d, e.side = bc.kind, int(f.loc[D] == o.VERTEX), f.dims[D], s
            v = bc.value
            if isinstance(v, o.BoundaryFunction):        # evaluated by the binding, uploaded as a value field
                v = o.boundary_value_field(g, f, bc, D, s)
            if isinstance(v, o.Field):
                pv = Pitched(v)
                keep.append(pv)
                ov = pv.opr()
                e.value, e.vp, e.vsy = 0.0, ov.p, ov.sy
            else:
                e.value, e.vp, e.vsy = (0.0 if v is None else float(v)), None, 0
    t = 0
    b.nt[0] = b.nt[1] = 1
    for a in range(g.nd):
        if a != D:
            b.nt[t] = g.n[a] + 3                          # remove_dim(dim, nvertices + 2), batch.jl:181
            t += 1
    assert (emul.bc_emul_run_f32 if f32 else emul.bc_emul_run)(C.byref(b)) == 0
    for s in range(2):
        if sides[s]:
            o.bc_side(g, D, s, ("field", sides[s]))
    for f, p in pit.values():
        got = p.dense(g.nd)
        same = same_bits(f.data, got)
        assert same.all(), f"dim {D} loc {f.loc}: {np.argwhere(~same)[:3]}"
    for p in keep + [p for _, p in pit.values()]:         # the padding past each row was never written
        body = p.flat[p.lead:p.lead + p.pitch * p.sd[1] * p.sd[2]].reshape(p.sd[2], p.sd[1], p.pitch)
        assert (body[:, :, p.sd[0]:] == 777.25).all() and (p.flat[:p.lead] == 777.25).all()


GRIDS = [((-1.0,), (2.0,), (9,)), ((-1.0, 0.5), (2.0, 1.7), (7, 5)), ((-5.0, -5.0, -5.0), (10.0, 9.0, 8.0), (6, 5, 4)),
         ((0.0, 0.0), (1.0, 3.0), (130, 3)), ((0.0, 0.0, 0.0), (1.0, 3.0, 2.0), (3, 131, 2))]      # > one 128-thread block


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS)
def test_every_rule_every_side_and_location(oracle, emul, origin, extent, n, dtype):
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(11)
    for loc in itertools.product((0, 1), repeat=nd):
        for D in range(nd):
            for mk in (o.Dirichlet, o.Neumann):
                for val in (None, 0.0, 1.75, -3.0e-3):
                    f = rnd_field(o, g, loc, rng)
                    run_bc_both(o, emul, g, D, ([(f, mk(val))], None))                 # side 1 alone
                    run_bc_both(o, emul, g, D, (None, [(f, mk(val))]))                 # side 2 alone
                    run_bc_both(o, emul, g, D, ([(f, mk(val))], [(f, mk(-2.5))]))      # both sides in one launch
                # different kinds on the two sides of one field
                f = rnd_field(o, g, loc, rng)
                other = o.Neumann if mk is o.Dirichlet else o.Dirichlet
                run_bc_both(o, emul, g, D, ([(f, mk(0.25))], [(f, other(-0.5))]))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS[1:3])
def test_field_valued_and_function_valued_conditions(oracle, emul, origin, extent, n, dtype):
    """value(bc, ...) = bc.value[remove_dim(dim, I)...] (first_order_boundary_condition.jl:38-40) and BoundaryFunction
    (boundary_function.jl:34-40), the latter evaluated host-side over the face range and uploaded."""
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(12)
    for loc in itertools.product((0, 1), repeat=nd):
        for D in range(nd):
            tg = o.transverse_grid(g, D)
            for mk in (o.Dirichlet, o.Neumann):
                for vloc in itertools.product((0, 1), repeat=nd - 1):
                    f = rnd_field(o, g, loc, rng)
                    v1, v2 = rnd_field(o, tg, vloc, rng), rnd_field(o, tg, vloc, rng)
                    run_bc_both(o, emul, g, D, ([(f, mk(v1))], [(f, mk(v2))]))
                f = rnd_field(o, g, loc, rng)
                cont = o.BoundaryFunction(lambda *x: 0.5 + sum((a + 1) * c for a, c in enumerate(x)))
                disc = o.BoundaryFunction(lambda grid, l, dim, *I: 0.125 * sum(I) - 0.5 * dim + l, discrete=True)
                run_bc_both(o, emul, g, D, ([(f, mk(cont))], [(f, mk(disc))]))


@pytest.mark.parametrize("dtype", DTYPES)
def test_full_batches_of_the_solvers(oracle, emul, dtype):
    """The batches the Stokes drivers pass (stokes_3d_inc_ve_T.jl: free slip = Dirichlet on the normal velocity, Neumann
    on the tangential ones and on Pr/T), 8 fields per side -- the widest batch one launch carries."""
    o = oracle
    g = o.Grid((-0.5, -0.5, -0.5), (1.0, 1.0, 1.0), (9, 7, 6), dtype=dtype)
    rng = np.random.default_rng(13)
    V = o.VectorField(g)
    tau = o.TensorField(g)
    for F in list(V.values()) + list(tau.values()):
        F.data[...] = rng.random(F.sdims) - 0.5
    T = rnd_field(o, g, 0, rng)
    for D in range(3):
        ax = "xyz"[D]
        lst = []
        for c, F in V.items():
            lst.append((F, o.Dirichlet() if c == ax else o.Neumann()))
        for c, F in tau.items():
            if len(lst) < MAXF - 1 and ax in c and c[0] != c[1]:
                lst.append((F, o.Dirichlet(0.0)))
        lst.append((T, o.Dirichlet(0.5) if D == 2 else o.Neumann(0.25)))
        while len(lst) < MAXF:
            lst.append((rnd_field(o, g, 1, rng), o.Neumann(-1.5)))
        assert len(lst) == MAXF
        right = [(F, o.Neumann(0.75) if bc.kind == o.DIRICHLET and F.loc[D] == o.CENTER else bc) for F, bc in lst]
        run_bc_both(o, emul, g, D, (lst, right))


@pytest.mark.parametrize("dtype", DTYPES)
def test_special_values_pass_through(oracle, emul, dtype):
    """NaN, infinities and signed zeros behave as in the reference's muladd rules."""
    o = oracle
    g = o.Grid((0.0, 0.0), (1.0, 1.0), (6, 5), dtype=dtype)
    rng = np.random.default_rng(14)
    for loc in itertools.product((0, 1), repeat=2):
        for D in range(2):
            for mk in (o.Dirichlet, o.Neumann):
                f = rnd_field(o, g, loc, rng)
                flat = f.data.reshape(-1, order="F")
                flat[::7], flat[3::11], flat[5::13], flat[1::17] = np.nan, -0.0, np.inf, 0.0
                run_bc_both(o, emul, g, D, ([(f, mk(-0.0))], [(f, mk(np.inf))]))


# ---------------------------------------------------------------------------------------------- halo slabs

def slab_batch(o, g, fields, pits, D, S, send):
    """make_slab_batch (bc.cu): slab index, transverse storage extents and offsets of the fields' slabs in one message."""
    f32 = g.dtype == np.float32
    b = (S32 if f32 else S64)[1]()
    b.n, b.dim, b.nd = len(fields), D, g.nd
    off = 0
    for q, (f, p) in enumerate(zip(fields, pits)):
        ov = 1 if f.loc[D] == o.VERTEX else 0
        d = f.dims[D]
        e = b.e[q]
        e.f = view_of(p)
        e.idx = ((1 + ov) if S == 0 else d - ov) if send else (0 if S == 0 else d + 1)
        ext = [f.sdims[a] for a in range(g.nd) if a != D] + [1, 1]
        e.e0, e.e1, e.off = ext[0], ext[1], off
        off += ext[0] * ext[1]
    return b, off


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS)
def test_pack_and_unpack_every_slab(oracle, emul, origin, extent, n, dtype):
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(15)
    run = emul.slab_emul_run_f32 if dtype == np.float32 else emul.slab_emul_run
    RP = C.POINTER(C.c_float if dtype == np.float32 else C.c_double)
    locs = list(itertools.product((0, 1), repeat=nd))
    for D in range(nd):
        for S in range(2):
            # one message carrying a field of every location (<= 8), as ExchangeBatch does (exchange_halo.jl:13-61)
            fields = [rnd_field(o, g, loc, rng) for loc in locs]
            pits = [Pitched(f) for f in fields]
            b, total = slab_batch(o, g, fields, pits, D, S, send=True)
            buf = np.full(total + 3, -7.5, dtype=dtype)
            assert run(C.byref(b), buf.ctypes.data_as(RP), 1) == 0
            want = np.concatenate([o.pack_send(f, D, S) for f in fields])
            assert want.size == total and (same_bits(buf[:total], want)).all() and (buf[total:] == -7.5).all()
            for f, p in zip(fields, pits):                       # packing reads only
                assert np.array_equal(f.data, p.dense(nd))
            # the neighbour unpacks the same message on its opposite side
            recv = [rnd_field(o, g, loc, rng) for loc in locs]
            rpit = [Pitched(f) for f in recv]
            rb, rtotal = slab_batch(o, g, recv, rpit, D, 1 - S, send=False)
            assert rtotal == total
            assert run(C.byref(rb), buf.ctypes.data_as(RP), 0) == 0
            off = 0
            for f, p in zip(recv, rpit):
                m = int(np.prod([f.sdims[a] for a in range(nd) if a != D], dtype=np.int64))
                o.unpack_recv(f, D, 1 - S, np.ascontiguousarray(want[off:off + m]))
                off += m
                same = same_bits(f.data, p.dense(nd))
                assert same.all(), f"unpack dim {D} side {1 - S} loc {f.loc}: {np.argwhere(~same)[:3]}"
                body = p.flat[p.lead:p.lead + p.pitch * p.sd[1] * p.sd[2]].reshape(p.sd[2], p.sd[1], p.pitch)
                assert (body[:, :, p.sd[0]:] == 777.25).all()
            assert (same_bits(buf[:total], want)).all()          # unpacking reads the message only


@pytest.mark.parametrize("dtype", DTYPES)
def test_exchange_between_two_ranks_equals_the_reference_rule(oracle, emul, dtype):
    """Send slab of rank 0's side 2 lands in rank 1's side-1 halo and vice versa (exchange_halo.jl:101-108): after the
    exchange a Vertex field's shared node and both ranks' halos hold the neighbour's values, bit for bit."""
    o = oracle
    g = o.Grid((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (5, 4, 3), dtype=dtype)
    rng = np.random.default_rng(16)
    run = emul.slab_emul_run_f32 if dtype == np.float32 else emul.slab_emul_run
    RP = C.POINTER(C.c_float if dtype == np.float32 else C.c_double)
    for D in range(3):
        for loc in ((0, 0, 0), (1, 1, 1), flip((0, 0, 0), D)):
            A, Bf = rnd_field(o, g, loc, rng), rnd_field(o, g, loc, rng)
            pa, pb = Pitched(A), Pitched(Bf)
            sb, total = slab_batch(o, g, [A], [pa], D, 1, send=True)
            msg_ab = np.empty(total, dtype=dtype)
            run(C.byref(sb), msg_ab.ctypes.data_as(RP), 1)
            sb, _ = slab_batch(o, g, [Bf], [pb], D, 0, send=True)
            msg_ba = np.empty(total, dtype=dtype)
            run(C.byref(sb), msg_ba.ctypes.data_as(RP), 1)
            rb, _ = slab_batch(o, g, [Bf], [pb], D, 0, send=False)
            run(C.byref(rb), msg_ab.ctypes.data_as(RP), 0)
            rb, _ = slab_batch(o, g, [A], [pa], D, 1, send=False)
            run(C.byref(rb), msg_ba.ctypes.data_as(RP), 0)
            a, b = pa.dense(3), pb.dense(3)
            d, ov = A.dims[D], int(loc[D] == o.VERTEX)
            sl = lambda i: tuple(slice(None) if x != D else i + 1 for x in range(3))      # logical i -> storage i+1
            assert np.array_equal(b[sl(0)], A.data[sl(d - ov)]) and np.array_equal(a[sl(d + 1)], Bf.data[sl(1 + ov)])


# ---------------------------------------------------------------------------------------------- all dimensions, one launch
def run_bc_all_both(o, emul, g, field_bcs, rev):
    """field_bcs: [(oracle field, {dim: (left bc | None, right bc | None)})].  The oracle applies the batch set the
    reference's way -- D = N..1, side 1 then side 2 (batch.jl:20-29) -- the emulation runs ONE k_bc_all launch
    (bc_all_point, bc_kernels.cuh) on PITCHED copies; whole padded arrays must agree bit for bit."""
    f32 = g.dtype == np.float32
    A = (S32 if f32 else S64)[2]
    b = A()
    b.nf, b.nd = len(field_bcs), g.nd
    for a in range(g.nd):
        b.n[a], b.spacing[a] = g.n[a], g.spacing[a]
    keep, pit = [], []
    for q, (f, per_dim) in enumerate(field_bcs):
        p = Pitched(f)
        pit.append((f, p))
        F = b.fld[q]
        F.f = view_of(p)
        for a in range(3):
            F.d[a] = f.dims[a] if a < g.nd else 1
            F.vertex[a] = int(a < g.nd and f.loc[a] == o.VERTEX)
            for s in range(2):
                F.r[a][s].kind = -1
        for D, pair in per_dim.items():
            for s, bc in enumerate(pair):
                if bc is None:
                    continue
                r = F.r[D][s]
                r.kind = bc.kind
                v = bc.value
                if isinstance(v, o.BoundaryFunction):
                    v = o.boundary_value_field(g, f, bc, D, s)
                if isinstance(v, o.Field):
                    pv = Pitched(v)
                    keep.append(pv)
                    ov = pv.opr()
                    r.value, r.vp, r.vsy = 0.0, ov.p, ov.sy
                else:
                    r.value, r.vp, r.vsy = (0.0 if v is None else float(v)), None, 0
    for q in range(b.nf):                                 # the faces that carry a condition: one slice of the launch grid each
        for D in range(g.nd):
            for sd in range(2):
                if b.fld[q].r[D][sd].kind >= 0:
                    b.act[b.nact] = (q * 3 + D) * 2 + sd
                    b.nact += 1
    assert (emul.bc_all_emul_run_f32 if f32 else emul.bc_all_emul_run)(C.byref(b), int(rev)) == 0
    for D in reversed(range(g.nd)):
        for s in range(2):
            lst = [(f, per_dim[D][s]) for f, per_dim in field_bcs if D in per_dim and per_dim[D][s] is not None]
            if lst:
                o.bc_side(g, D, s, ("field", lst))
    for f, p in pit:
        same = same_bits(f.data, p.dense(g.nd))
        assert same.all(), f"loc {f.loc}: {np.argwhere(~same)[:4]}"
    for p in keep + [p for _, p in pit]:
        body = p.flat[p.lead:p.lead + p.pitch * p.sd[1] * p.sd[2]].reshape(p.sd[2], p.sd[1], p.pitch)
        assert (body[:, :, p.sd[0]:] == 777.25).all() and (p.flat[:p.lead] == 777.25).all()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("origin,extent,n", GRIDS + [((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (1, 1, 1)), ((0.0, 0.0), (1.0, 2.0), (1, 2))])
def test_all_dimensions_in_one_launch_equal_the_sequential_order(oracle, emul, origin, extent, n, dtype):
    """Every combination that makes the D = N..1 order observable: edges and corners where a later dimension reads what an
    earlier one wrote, Dirichlet nodes of Vertex fields that a later Neumann face copies, missing sides / dimensions,
    valued conditions whose value is looked up at the MOVED cell's transverse index -- for every staggered location."""
    o, nd = oracle, len(n)
    g = o.Grid(origin, extent, n, dtype=dtype)
    rng = np.random.default_rng(21)
    kinds = [o.Dirichlet, o.Neumann]
    for loc in itertools.product((0, 1), repeat=nd):
        for trial in range(6):
            per_dim = {}
            for D in range(nd):
                if trial == 5 and D == nd - 1 and nd > 1:
                    continue                                        # a dimension without any condition
                pair = []
                for s in range(2):
                    mk = kinds[(trial + D + s + loc[D]) % 2]
                    val = [None, 0.0, 1.75, -3.0e-3, 0.5, -1.25][(trial + 2 * D + s) % 6]
                    pair.append(None if (trial == 4 and s == D % 2) else mk(val))
                per_dim[D] = tuple(pair)
            f = rnd_field(o, g, loc, rng)
            run_bc_all_both(o, emul, g, [(f, per_dim)], rev=trial % 2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_all_dimensions_in_one_launch_solver_batches_and_valued_conditions(oracle, emul, dtype):
    o = oracle
    g = o.Grid((-0.5, -0.5, -0.5), (1.0, 1.0, 1.0), (9, 7, 6), dtype=dtype)
    rng = np.random.default_rng(22)
    V = o.VectorField(g)
    for F in V.values():
        F.data[...] = rng.random(F.sdims) - 0.5
    T = rnd_field(o, g, 0, rng)
    # the velocity batch of the Stokes drivers (stokes_3d_inc_ve_T.jl:130-132) + T Neumann (:133) in ONE launch
    fb = [(F, {D: ((o.Dirichlet() if c == "xyz"[D] else o.Neumann()),) * 2 for D in range(3)}) for c, F in V.items()]
    fb.append((T, {D: (o.Neumann(), o.Neumann()) for D in range(3)}))
    for rev in (0, 1):
        run_bc_all_both(o, emul, g, fb, rev)
    # Field- and function-valued conditions on every dimension of one field
    for loc in itertools.product((0, 1), repeat=3):
        f = rnd_field(o, g, loc, rng)
        per_dim = {}
        for D in range(3):
            tg = o.transverse_grid(g, D)
            v1 = rnd_field(o, tg, tuple((loc[a] + D) % 2 for a in range(3) if a != D), rng)
            cont = o.BoundaryFunction(lambda *x: 0.5 + sum((a + 1) * c for a, c in enumerate(x)))
            per_dim[D] = (o.Dirichlet(v1) if D != 1 else o.Neumann(v1), o.Neumann(cont) if D != 2 else o.Dirichlet(cont))
        run_bc_all_both(o, emul, g, [(f, per_dim)], rev=sum(loc) % 2)
