"""
oracle.py -- Python/ctypes front end of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY.

Restates, independently of the product's host layer (chmy.jl_b200/), the *control flow* of the reference
(PTsolvers/Chmy.jl v0.1.25; citations are file:line under /root/reference) around the C arithmetic in
chmy_oracle.c:

  * Launcher region algebra and the inner/outer launch order   src/KernelLaunch.jl:40-87,105-183
  * batch() normalisation and the bc! application order          src/BoundaryConditions/batch.jl:20-29,72-155
  * CartesianTopology (MPI_Dims_create / Cart_* semantics)       src/Distributed/topology.jl:26-41
  * distributed sub-grid                                         src/Distributed/distributed_grid.jl:1-36
  * exchange_halo! ordering, simulated in-process over all ranks src/Distributed/exchange_halo.jl:13-84

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The reference
cannot be run in this image (no Julia): the oracle is pinned on the reference's own known-answer tests
(tests/test_oracle_golden.py); halo exchange, launch splitting and the solver drivers are "parity unpinned"
and are checked through invariants instead.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field as dc_field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHMY_ORACLE_SANITIZE=1 (tests/test_emulation_sanitizers.py's child process, which preloads the sanitizer runtimes) loads
# AddressSanitizer + UBSan builds of the same source instead
_SAN = "_san" if os.environ.get("CHMY_ORACLE_SANITIZE") == "1" else ""
_LIB_PATH = os.path.join(_HERE, f"libchmy_oracle{_SAN}.so")            # og_real = double (the solvers' element type)
_LIB_PATH_F32 = os.path.join(_HERE, f"libchmy_oracle_f32{_SAN}.so")    # og_real = float  (the Float32 rows of the reference's tests)

CENTER, VERTEX = 0, 1
BOUNDED, CONNECTED = 0, 1
DIRICHLET, NEUMANN = 0, 1
AXES = ("x", "y", "z")


def _native_path() -> str:
    """CHMY_ORACLE_NATIVE=1 (bench.py's CPU arm): the Float64 build compiled -O3 -march=native ON THIS HOST, named after
    the host's CPU (model + flags) so that a copy built on another machine is never loaded."""
    import hashlib
    try:
        info = open("/proc/cpuinfo").read()
        key = "".join(l for l in info.splitlines()[:30] if l.startswith(("model name", "flags")))
    except OSError:
        key = "unknown"
    return os.path.join(_HERE, f"libchmy_oracle_native_{hashlib.sha1(key.encode()).hexdigest()[:10]}.so")


def build(force: bool = False) -> str:
    """Compile chmy_oracle.c with the committed Makefile (gcc, -ffp-contract=off): the Float64 and the Float32 build."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("chmy_oracle.c", "chmy_oracle.h", "Makefile"))
    for path in (_LIB_PATH, _LIB_PATH_F32):
        if force or not os.path.exists(path) or os.path.getmtime(path) < src_m:
            subprocess.run(["make", "-C", _HERE, "-B", os.path.basename(path)], check=True, capture_output=True)
    if os.environ.get("CHMY_ORACLE_NATIVE") == "1":
        path = _native_path()
        if force or not os.path.exists(path) or os.path.getmtime(path) < src_m:
            subprocess.run(["make", "-C", _HERE, "-B", "native", "OUT=" + path], check=True, capture_output=True)
    return _LIB_PATH


class _Binding:
    """ctypes view of one build of chmy_oracle.c: struct layouts and prototypes in its element type."""

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float64), np.dtype(np.float32))
        R = self.R = C.c_double if self.dtype == np.float64 else C.c_float
        path = _LIB_PATH if self.dtype == np.float64 else _LIB_PATH_F32
        if self.dtype == np.float64 and os.environ.get("CHMY_ORACLE_NATIVE") == "1":
            path = _native_path()

        class CGrid(C.Structure):
            _fields_ = [("nd", C.c_int32), ("n", C.c_int64 * 3), ("origin", R * 3), ("extent", R * 3),
                        ("spacing", R * 3), ("inv_spacing", R * 3), ("conn", (C.c_int32 * 2) * 3)]

        class CField(C.Structure):
            _fields_ = [("nd", C.c_int32), ("loc", C.c_int32 * 3), ("d", C.c_int64 * 3), ("sd", C.c_int64 * 3),
                        ("o", C.c_int64 * 3), ("data", C.POINTER(R))]

        class CIncl(C.Structure):
            _fields_ = [("active", C.c_int32), ("loc", C.c_int32 * 3), ("c0", R * 3), ("r", R), ("inn", R), ("out", R)]

        self.CGrid, self.CField, self.CIncl = CGrid, CField, CIncl
        try:
            build()
            L = C.CDLL(path)
        except OSError:
            build(force=True)
            L = C.CDLL(path)
        self.lib = L
        P = C.POINTER
        L.og_real_bytes.restype = C.c_int
        assert L.og_real_bytes() == C.sizeof(R)
        L.og_grid_init.argtypes = [P(CGrid), C.c_int, P(C.c_int64), P(R), P(R)]
        L.og_coord.argtypes = [P(CGrid), C.c_int, C.c_int, C.c_int64]
        L.og_coord.restype = R
        L.og_field_init.argtypes = [P(CField), P(CGrid), P(C.c_int32), P(R)]
        L.og_field_storage_len.argtypes = [P(CGrid), P(C.c_int32)]
        L.og_field_storage_len.restype = C.c_int64
        L.og_set_inclusion.argtypes = [P(CGrid), P(CField), P(CIncl)]
        L.og_maxabs_interior.argtypes = [P(CField)]
        L.og_maxabs_interior.restype = R
        box = [P(C.c_int64), P(C.c_int64)]
        FP, FPP = P(CField), P(P(CField))
        L.og_compute_q.argtypes = [P(CGrid), FP, FP, FP, R] + box
        L.og_update_C.argtypes = [P(CGrid), FP, FP, FP, R] + box
        L.og_update_old.argtypes = [P(CGrid), C.c_int, FPP, FPP] + box
        for nm in ("og_update_stress2", "og_update_stress3"):
            getattr(L, nm).argtypes = [P(CGrid), FPP, FP, FP, FPP, FPP] + [R] * 6 + box
        for nm in ("og_update_velocity2", "og_update_velocity3"):
            getattr(L, nm).argtypes = [P(CGrid), FPP, FPP, FP, FPP, FP, P(CIncl), R, R] + box
        L.og_update_thermal_flux.argtypes = [P(CGrid), FPP, FP, FPP, R] + box
        L.og_update_thermal.argtypes = [P(CGrid), FP, FP, FPP, R] + box
        L.og_bc_apply.argtypes = [P(CGrid), FP, C.c_int, C.c_int, C.c_int, R]
        L.og_bc_apply_field.argtypes = [P(CGrid), FP, C.c_int, C.c_int, C.c_int, R, FP]
        L.og_slab_len.argtypes = [FP, C.c_int]
        L.og_slab_len.restype = C.c_int64
        L.og_pack_send.argtypes = [FP, C.c_int, C.c_int, P(R)]
        L.og_unpack_recv.argtypes = [FP, C.c_int, C.c_int, P(R)]
        for nm in ("og_partial", "og_partial2"):
            getattr(L, nm).argtypes = [P(CGrid), FP, C.c_int, C.c_int64, C.c_int64, C.c_int64]
            getattr(L, nm).restype = R
        for nm in ("og_lerp", "og_hlerp"):
            getattr(L, nm).argtypes = [P(CGrid), FP, P(C.c_int32), C.c_int64, C.c_int64, C.c_int64]
            getattr(L, nm).restype = R
        L.og_dkd.argtypes = [P(CGrid), FP, FP, C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.og_dkd.restype = R
        L.og_apply_operator.argtypes = [P(CGrid), C.c_int, C.c_int, FPP, FPP, FP] + box
        L.og_num_threads.restype = C.c_int
        L.og_set_num_threads.argtypes = [C.c_int]
        L.og_set_num_threads.restype = None


_bindings: Dict[str, _Binding] = {}


def binding(dtype=np.float64) -> _Binding:
    key = np.dtype(dtype).name
    if key not in _bindings:
        _bindings[key] = _Binding(dtype)
    return _bindings[key]


def lib():
    """The Float64 build (the element type of every solver on this path)."""
    return binding(np.float64).lib


def num_threads() -> int:
    return int(lib().og_num_threads())


def set_num_threads(n: int) -> int:
    lib().og_set_num_threads(int(n))
    return num_threads()


# ------------------------------------------------------------------------------------------------ grid

def expand_loc(nd: int, loc) -> Tuple[int, ...]:
    """src/Grids/Grids.jl expand_loc: a single Location repeats over all dims."""
    if isinstance(loc, int):
        return (loc,) * nd
    assert len(loc) == nd
    return tuple(int(l) for l in loc)


class Grid:
    """UniformGrid(arch; origin, extent, dims, topology)  -- src/Grids/structured_grid.jl:27-39."""

    def __init__(self, origin, extent, dims, conn=None, dtype=np.float64):
        nd = len(dims)
        self.nd = nd
        self.B = binding(dtype)                 # eltype(grid): Float64 | Float32 (test/common.jl:9)
        self.dtype = self.B.dtype
        self.c = self.B.CGrid()
        n = (C.c_int64 * 3)(*([int(x) for x in dims] + [1] * (3 - nd)))
        o = (self.B.R * 3)(*([float(x) for x in origin] + [0.0] * (3 - nd)))
        e = (self.B.R * 3)(*([float(x) for x in extent] + [0.0] * (3 - nd)))
        self.B.lib.og_grid_init(C.byref(self.c), nd, n, o, e)
        self.conn = [[BOUNDED, BOUNDED] for _ in range(nd)] if conn is None else [list(c) for c in conn]
        for d in range(nd):
            for s in range(2):
                self.c.conn[d][s] = self.conn[d][s]

    @property
    def n(self):
        return tuple(int(self.c.n[d]) for d in range(self.nd))

    @property
    def origin(self):
        return tuple(self.c.origin[d] for d in range(self.nd))

    @property
    def extent(self):
        return tuple(self.c.extent[d] for d in range(self.nd))

    @property
    def spacing(self):
        return tuple(self.c.spacing[d] for d in range(self.nd))

    @property
    def inv_spacing(self):
        return tuple(self.c.inv_spacing[d] for d in range(self.nd))

    def size(self, loc) -> Tuple[int, ...]:
        """size(grid, loc) -- structured_grid.jl:49 ; abstract_axis.jl:10-13."""
        loc = expand_loc(self.nd, loc)
        return tuple(self.n[d] + (1 if loc[d] == VERTEX else 0) for d in range(self.nd))

    def coord(self, dim: int, loc: int, i: int) -> float:
        """coord(grid, loc, Dim(dim+1), i) with 1-based i -- uniform_axis.jl:18-19."""
        return float(self.B.lib.og_coord(C.byref(self.c), dim, loc, int(i)))

    def coords(self, dim: int, loc: int) -> np.ndarray:
        d = self.n[dim] + (1 if loc == VERTEX else 0)
        return np.array([self.coord(dim, loc, i) for i in range(1, d + 1)], dtype=self.dtype)

    # abstract_axis.jl:23-36 / uniform_axis.jl:21-25
    def origin_at(self, dim, loc):
        return self.coord(dim, loc, 1) if loc == CENTER else self.c.origin[dim]

    def extent_at(self, dim, loc):
        T = self.dtype.type                   # extent(ax, ::Center) = extent - spacing in eltype(ax) (uniform_axis.jl:25)
        return float(self.c.extent[dim]) if loc == VERTEX else float(T(self.c.extent[dim]) - T(self.c.spacing[dim]))

    def bounds(self, dim, loc):
        T = self.dtype.type
        o = self.origin_at(dim, loc)
        return (o, float(T(o) + T(self.extent_at(dim, loc))))


class Field:
    """Field(backend, grid, loc; halo=1) -- src/Fields/field.jl:56-62 (zero-initialised, dims+4 per dim)."""

    def __init__(self, grid: Grid, loc=CENTER):
        self.grid = grid
        self.nd = grid.nd
        self.loc = expand_loc(grid.nd, loc)
        self.dims = grid.size(self.loc)
        self.sdims = tuple(d + 4 for d in self.dims)
        self.B, self.dtype = grid.B, grid.dtype
        self.data = np.zeros(self.sdims, dtype=self.dtype, order="F")
        self.c = self.B.CField()
        locs = (C.c_int32 * 3)(*(list(self.loc) + [0] * (3 - self.nd)))
        self.B.lib.og_field_init(C.byref(self.c), C.byref(grid.c), locs,
                                 self.data.ctypes.data_as(C.POINTER(self.B.R)))

    def interior(self, with_halo: bool = False) -> np.ndarray:
        """interior(f; with_halo) -- field.jl:33-37 (a view)."""
        h = 1 if with_halo else 0
        return self.data[tuple(slice(2 - h, 2 + d + h) for d in self.dims)]

    def set(self, val):
        """set!(f, val::Number) / set!(f, A::AbstractArray) -- field.jl:87-98 (interior only)."""
        self.interior()[...] = val

    def set_fun(self, fun, *params):
        """set!(f, grid, fun; parameters) continuous -- field.jl:121-124,131-142."""
        cs = [self.grid.coords(d, self.loc[d]) for d in range(self.nd)]
        mesh = np.meshgrid(*cs, indexing="ij")
        self.interior()[...] = fun(*mesh, *params)

    def maxabs(self) -> float:
        return float(self.B.lib.og_maxabs_interior(C.byref(self.c)))

    def at(self, *I):
        """logical (reference, 1-based) indexing f[I...] -- field.jl:18."""
        return self.data[tuple(i + 1 for i in I)]


def VectorField(grid: Grid) -> Dict[str, Field]:
    """field.jl:148-169 : component D is Vertex along D, Center elsewhere."""
    return {AXES[D]: Field(grid, tuple(VERTEX if i == D else CENTER for i in range(grid.nd))) for D in range(grid.nd)}


def TensorField(grid: Grid) -> Dict[str, Field]:
    """field.jl:182-206."""
    if grid.nd == 2:
        return {"xx": Field(grid, CENTER), "yy": Field(grid, CENTER), "xy": Field(grid, VERTEX)}
    return {"xx": Field(grid, CENTER), "yy": Field(grid, CENTER), "zz": Field(grid, CENTER),
            "xy": Field(grid, (VERTEX, VERTEX, CENTER)), "xz": Field(grid, (VERTEX, CENTER, VERTEX)),
            "yz": Field(grid, (CENTER, VERTEX, VERTEX))}


@dataclass
class Inclusion:
    """FunctionField(init_incl, grid, loc; parameters=(x0,y0[,z0],r,in,out)) -- function_field.jl:12-59."""
    loc: Tuple[int, ...]
    c0: Tuple[float, ...]
    r: float
    inn: float
    out: float

    def cstruct(self, B: Optional[_Binding] = None):
        s = (B or binding()).CIncl()
        s.active = 1
        for d in range(len(self.loc)):
            s.loc[d] = self.loc[d]
            s.c0[d] = self.c0[d]
        s.r, s.inn, s.out = self.r, self.inn, self.out
        return s


def set_inclusion(f: Field, inc: Inclusion):
    cs = inc.cstruct(f.B)
    f.B.lib.og_set_inclusion(C.byref(f.grid.c), C.byref(f.c), C.byref(cs))


# ------------------------------------------------------------------------------------------------ BCs

@dataclass(frozen=True)
class BC:
    kind: int
    value: object = None     # None == `nothing` -> zero(eltype) (first_order_boundary_condition.jl:34); Number (:36);
                             # lower-dimensional Field read at remove_dim(dim, I) (:38-40); BoundaryFunction


class BoundaryFunction:
    """boundary_function.jl:6-44,69-72.  bf(grid, loc, dim, I...) with ONE location for all axes (the location of
    the field along the boundary dim, batch.jl:174; coord(grid, loc::Location, I...) structured_grid.jl:103-106):
      continuous: fun(reduce(dim, coord(grid, loc, I...))..., params...)      (:34-36)
      discrete  : fun(grid, loc, dim, reduce(dim, I)..., params...)           (:38-40)
    reduce = remove_dim(dim, .) unless reduce_dims=false (:31-32).  dim is 1-based; loc is CENTER | VERTEX."""

    def __init__(self, fun, discrete=False, parameters=None, reduce_dims=True):
        self.fun, self.discrete, self.parameters, self.reduce_dims = fun, discrete, parameters, reduce_dims

    def _params(self):
        if self.parameters is None:
            return ()
        return tuple(self.parameters) if isinstance(self.parameters, (tuple, list)) else (self.parameters,)

    def __call__(self, grid, loc, dim, *I):
        red = (lambda t: tuple(x for a, x in enumerate(t, start=1) if a != dim)) if self.reduce_dims else (lambda t: tuple(t))
        if self.discrete:
            return self.fun(grid, loc, dim, *red(I), *self._params())
        x = tuple(grid.coord(a, loc, i) for a, i in enumerate(I))
        return self.fun(*red(x), *self._params())


def Dirichlet(value=None) -> BC:
    return BC(DIRICHLET, value)


def Neumann(value=None) -> BC:
    return BC(NEUMANN, value)


def _regularise_bc(nd, spec):
    """batch.jl:126-130 : a single BC applies to every dim and side; a named tuple fills only the named axes."""
    if isinstance(spec, BC):
        return [(spec, spec) for _ in range(nd)]
    out = [(None, None) for _ in range(nd)]
    for name, v in spec.items():
        D = AXES.index(name)
        out[D] = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    return out


def _regularise_exchange(nd, exchange):
    """batch.jl:135-141."""
    if exchange is None:
        return [None] * nd
    if isinstance(exchange, Field):
        return [(exchange,)] * nd
    if isinstance(exchange, dict):
        out = [None] * nd
        for name, f in exchange.items():
            out[AXES.index(name)] = (f,) if isinstance(f, Field) else tuple(f)
        return out
    return [tuple(exchange)] * nd


def batch(grid: Grid, *field_bcs, exchange=None):
    """batch(grid, f => bc...; exchange) -> BatchSet  -- batch.jl:72-116.
    Returns [ (left, right) per dim ], each side ('empty',) | ('field', [(f, bc)...]) | ('exchange', (f...))."""
    nd = grid.nd
    fields = [fb[0] for fb in field_bcs]
    bcs = [_regularise_bc(nd, fb[1]) for fb in field_bcs]
    exch = _regularise_exchange(nd, exchange)
    out = []
    for D in range(nd):
        sides = []
        for S in range(2):
            if grid.conn[D][S] == CONNECTED:                       # batch.jl:98-101
                e = exch[D]
                sides.append(("exchange", tuple(e)) if e else ("empty",))
            else:                                                  # batch.jl:103-105 + prune :151-155
                fb = [(f, b[D][S]) for f, b in zip(fields, bcs) if b[D][S] is not None]
                sides.append(("field", fb) if fb else ("empty",))
        out.append(tuple(sides))
    return out


def bc_side(grid: Grid, D: int, S: int, b):
    """bc!(side, dim, arch, grid, ::FieldBatch) -- batch.jl:163-184 (fields in batch order)."""
    if b[0] == "field":
        for f, bc in b[1]:
            v = bc.value
            if isinstance(v, BoundaryFunction):
                v = boundary_value_field(grid, f, bc, D, S)
            if isinstance(v, Field):
                grid.B.lib.og_bc_apply_field(C.byref(grid.c), C.byref(f.c), D, S, bc.kind, 0.0, C.byref(v.c))
            else:
                grid.B.lib.og_bc_apply(C.byref(grid.c), C.byref(f.c), D, S, bc.kind, 0.0 if v is None else float(v))


def transverse_grid(grid: Grid, D: int) -> Grid:
    keep = [a for a in range(grid.nd) if a != D]
    return Grid([grid.origin[a] for a in keep], [grid.extent[a] for a in keep], [grid.n[a] for a in keep], dtype=grid.dtype)


def boundary_value_field(grid: Grid, f: Field, bc: BC, D: int, S: int) -> Field:
    """The values value(bc, grid, loc, dim, I_f...) takes over the face range 0..n_t+2 (batch.jl:159-184), with the
    (loc, index along D) each rule passes (first_order_boundary_condition.jl:42-84), as an (N-1)-dim Field."""
    d = f.dims[D]
    if bc.kind == DIRICHLET and f.loc[D] == VERTEX:
        loc, idx = VERTEX, (1 if S == 0 else d)
    elif bc.kind == DIRICHLET:
        loc, idx = CENTER, (0 if S == 0 else d + 1)
    else:
        loc, idx = 1 - f.loc[D], (0 if S == 0 else d + 1)
    tg = transverse_grid(grid, D)
    vf = Field(tg, VERTEX)
    ext = [n + 3 for n in tg.n]
    for J in np.ndindex(*ext):
        I = list(int(j) for j in J)
        I.insert(D, idx)
        vf.data[tuple(j + 1 for j in J)] = bc.value(grid, loc, D + 1, *I)
    return vf


# ------------------------------------------------------------------------------------------------ topology

def dims_create(nprocs: int, dims: Sequence[int]) -> Tuple[int, ...]:
    """MPI_Dims_create semantics used by topology.jl:28: fill zero entries with a balanced factorisation of
    nprocs / prod(fixed), as close to each other as possible, in non-increasing order."""
    dims = list(dims)
    fixed = 1
    for d in dims:
        if d > 0:
            fixed *= d
    assert nprocs % fixed == 0
    rest = nprocs // fixed
    free = [i for i, d in enumerate(dims) if d == 0]
    if not free:
        assert rest == 1
        return tuple(dims)
    primes = []
    p, m = 2, rest
    while p * p <= m:
        while m % p == 0:
            primes.append(p)
            m //= p
        p += 1
    if m > 1:
        primes.append(m)
    vals = [1] * len(free)
    for pr in sorted(primes, reverse=True):
        vals[vals.index(min(vals))] *= pr
    vals.sort(reverse=True)
    for i, v in zip(free, vals):
        dims[i] = v
    return tuple(dims)


@dataclass
class Topology:
    """CartesianTopology -- topology.jl:6-41 : MPI_Cart_create (non-periodic, row-major ranks), Cart_coords,
    Cart_shift(dim, 1) -> (left, right), PROC_NULL (= -1 here) at the edges."""
    nprocs: int
    dims: Tuple[int, ...]
    rank: int
    coords: Tuple[int, ...] = dc_field(init=False)
    neighbors: List[Tuple[int, int]] = dc_field(init=False)

    def __post_init__(self):
        self.coords = self.rank_to_coords(self.rank)
        nb = []
        for D in range(len(self.dims)):
            pair = []
            for delta in (-1, +1):
                c = list(self.coords)
                c[D] += delta
                pair.append(self.coords_to_rank(c) if 0 <= c[D] < self.dims[D] else -1)
            nb.append(tuple(pair))
        self.neighbors = nb

    def rank_to_coords(self, r):
        c = []
        for d in reversed(self.dims):
            c.append(r % d)
            r //= d
        return tuple(reversed(c))

    def coords_to_rank(self, c):
        r = 0
        for ci, d in zip(c, self.dims):
            r = r * d + ci
        return r


def local_grid(global_origin, global_extent, global_dims, topo: Topology, dtype=np.float64) -> Grid:
    """StructuredGrid{C}(arch::DistributedArchitecture, axes...) -- distributed_grid.jl:1-36.
    local_n = cld(global_n, dims); offset = coords*local_n; origin = vertex(ax, offset+1);
    extent = spacing*local_n; then UniformAxis recomputes spacing = extent/local_n."""
    nd = len(global_dims)
    g = Grid(global_origin, global_extent, global_dims, dtype=dtype)
    ln = [-(-global_dims[d] // topo.dims[d]) for d in range(nd)]
    off = [topo.coords[d] * ln[d] for d in range(nd)]
    new_origin = [g.coord(d, VERTEX, off[d] + 1) for d in range(nd)]
    new_extent = [g.c.spacing[d] * ln[d] for d in range(nd)]
    conn = [[CONNECTED if topo.neighbors[d][s] >= 0 else BOUNDED for s in range(2)] for d in range(nd)]
    return Grid(new_origin, new_extent, ln, conn, dtype=dtype)


# ------------------------------------------------------------------------------------------------ halo exchange

def pack_send(f: Field, D: int, S: int) -> np.ndarray:
    n = int(f.B.lib.og_slab_len(C.byref(f.c), D))
    buf = np.empty(n, dtype=f.dtype)
    f.B.lib.og_pack_send(C.byref(f.c), D, S, buf.ctypes.data_as(C.POINTER(f.B.R)))
    return buf


def unpack_recv(f: Field, D: int, S: int, buf: np.ndarray):
    assert buf.size == int(f.B.lib.og_slab_len(C.byref(f.c), D)) and buf.dtype == f.dtype
    f.B.lib.og_unpack_recv(C.byref(f.c), D, S, buf.ctypes.data_as(C.POINTER(f.B.R)))


# ------------------------------------------------------------------------------------------------ launcher

def _box(nd, lo, hi):
    lo3 = (C.c_int64 * 3)(*(list(lo) + [0] * (3 - nd)))
    hi3 = (C.c_int64 * 3)(*(list(hi) + [0] * (3 - nd)))
    return lo3, hi3


def _fparr(fs: Sequence[Field]):
    arr = (C.POINTER(fs[0].B.CField) * len(fs))(*[C.pointer(f.c) for f in fs])
    return arr


def _names(nd):
    return (("xx", "yy", "xy"), ("x", "y")) if nd == 2 else (("xx", "yy", "zz", "xy", "xz", "yz"), ("x", "y", "z"))


# op bodies: (grid, args, lo, hi) -> None.  Argument order follows the reference kernels' signatures.  Every op runs in the
# element type of its grid (Float64 | Float32, test/common.jl:9): scalars are converted to that type, Float64 literals inside
# the kernels promote as in Julia (chmy_oracle.h: og_wide).
def compute_q(g, args, lo, hi):
    q, Cf, chi = args
    g.B.lib.og_compute_q(C.byref(g.c), C.byref(q["x"].c), C.byref(q["y"].c), C.byref(Cf.c), chi, *_box(g.nd, lo, hi))


def update_C(g, args, lo, hi):
    Cf, q, dt = args
    g.B.lib.og_update_C(C.byref(g.c), C.byref(Cf.c), C.byref(q["x"].c), C.byref(q["y"].c), dt, *_box(g.nd, lo, hi))


def update_old(g, args, lo, hi):
    T, tau, T_old, tau_old = args
    tn, _ = _names(g.nd)
    dst = [T_old] + [tau_old[c] for c in tn]
    src = [T] + [tau[c] for c in tn]
    g.B.lib.og_update_old(C.byref(g.c), len(dst), _fparr(dst), _fparr(src), *_box(g.nd, lo, hi))


def update_stress(g, args, lo, hi):
    tau, Pr, divV, V, tau_old, eta, eta_ve, G, dt, dtau_Pr, dtau_r = args
    tn, vn = _names(g.nd)
    fn = g.B.lib.og_update_stress2 if g.nd == 2 else g.B.lib.og_update_stress3
    fn(C.byref(g.c), _fparr([tau[c] for c in tn]), C.byref(Pr.c), C.byref(divV.c), _fparr([V[c] for c in vn]),
       _fparr([tau_old[c] for c in tn]), eta, eta_ve, G, dt, dtau_Pr, dtau_r, *_box(g.nd, lo, hi))


def update_velocity(g, args, lo, hi):
    V, rV, Pr, tau, rhog, eta_ve, nudtau = args
    tn, vn = _names(g.nd)
    fn = g.B.lib.og_update_velocity2 if g.nd == 2 else g.B.lib.og_update_velocity3
    if isinstance(rhog, Inclusion):
        inc, fld = rhog.cstruct(g.B), None
    else:
        inc, fld = None, rhog
    fn(C.byref(g.c), _fparr([V[c] for c in vn]), _fparr([rV[c] for c in vn]), C.byref(Pr.c),
       _fparr([tau[c] for c in tn]), C.byref(fld.c) if fld is not None else None,
       C.byref(inc) if inc is not None else None, eta_ve, nudtau, *_box(g.nd, lo, hi))


def update_thermal_flux(g, args, lo, hi):
    qT, T, V, lam = args
    _, vn = _names(g.nd)
    g.B.lib.og_update_thermal_flux(C.byref(g.c), _fparr([qT[c] for c in vn]), C.byref(T.c),
                                 _fparr([V[c] for c in vn]), lam, *_box(g.nd, lo, hi))


def update_thermal(g, args, lo, hi):
    T, T_old, qT, dt = args
    _, vn = _names(g.nd)
    g.B.lib.og_update_thermal(C.byref(g.c), C.byref(T.c), C.byref(T_old.c), _fparr([qT[c] for c in vn]), dt,
                            *_box(g.nd, lo, hi))


class Launcher:
    """Launcher(arch, grid; outer_width) -- src/KernelLaunch.jl:21-87.  Work-item J in 1..worksize maps to the
    logical index I = J + Offset(-1) + region offset (KernelLaunch.jl:109,163,172)."""

    def __init__(self, grid: Grid, outer_width=None):
        self.nd = grid.nd
        self.worksize = tuple(n + 2 for n in grid.n)                      # :41
        self.outer_width = None if outer_width is None else tuple(outer_width)

    def regions(self):
        """[(name, lo, hi)] boxes in logical indices, in the reference's launch order."""
        ws, ow, N = self.worksize, self.outer_width, self.nd
        if ow is None:
            return [("full", tuple(0 for _ in ws), tuple(w - 1 for w in ws))]
        regs = []
        inner_ws = tuple(w - 2 * o for w, o in zip(ws, ow))               # :60
        inner_off = ow                                                    # :61
        regs.append(("inner", tuple(inner_off), tuple(o + w - 1 for o, w in zip(inner_off, inner_ws))))
        for D in reversed(range(N)):                                      # :166-168
            for S in range(2):
                size = tuple(ws[I] if I < D else ow[I] if I == D else ws[I] - 2 * ow[I] for I in range(N))   # :63-74
                off = tuple(0 if I < D else (0 if S == 0 else ws[I] - ow[I]) if I == D else ow[I] for I in range(N))  # :76-87
                regs.append((f"outer{D}{S}", off, tuple(o + s - 1 for o, s in zip(off, size))))
        return regs


def launch_world(launchers: Sequence[Launcher], grids: Sequence[Grid], op, args: Sequence, bcs: Sequence = None,
                 topos: Sequence[Topology] = None):
    """launcher(arch, grid, op => args; bc) executed for every rank of an in-process world in lock step.
    Literal order of KernelLaunch.jl:152-183: without bc or without outer_width one full-range kernel (then bc!);
    otherwise inner kernel, then for D = N..1: the two outer slabs of D followed by the (D, side) batches; sides of
    one dim are independent, dims are strictly sequential (:177-178)."""
    R = len(grids)
    nd = grids[0].nd
    split = bcs is not None and launchers[0].outer_width is not None
    if not split:
        for r in range(R):
            L = launchers[r]
            op(grids[r], args[r], tuple(0 for _ in L.worksize), tuple(w - 1 for w in L.worksize))
        if bcs is not None:
            bc_world(grids, bcs, topos)
        return
    regs = [L.regions() for L in launchers]
    for r in range(R):
        _, lo, hi = regs[r][0]
        if all(h >= l for l, h in zip(lo, hi)):
            op(grids[r], args[r], lo, hi)
    idx = 1
    for D in reversed(range(nd)):
        for S in range(2):
            for r in range(R):
                name, lo, hi = regs[r][idx]
                assert name == f"outer{D}{S}"
                if all(h >= l for l, h in zip(lo, hi)):          # zero-width slabs are legal
                    op(grids[r], args[r], lo, hi)
            idx += 1
        bc_dim_world(D, grids, bcs, topos)


def bc_dim_world(D: int, grids: Sequence[Grid], bcs: Sequence, topos: Sequence[Topology] = None):
    """Both sides of dim D on every rank.  FieldBatch sides: bc_kernel! (batch.jl:163-184).  ExchangeBatch sides:
    exchange_halo! (exchange_halo.jl:13-61,101-108): per field, send slab -> the neighbour's recv slab on the
    opposite side; all packs complete before any unpack (pack + Isend, then unpack as receives land)."""
    R = len(grids)
    for S in range(2):
        for r in range(R):
            bc_side(grids[r], D, S, bcs[r][D][S])
    mail = {}
    for r in range(R):
        for S in range(2):
            b = bcs[r][D][S]
            if b[0] == "exchange":
                nb = topos[r].neighbors[D][S]
                assert nb >= 0, "no neighbor to communicate"           # exchange_halo.jl:19
                mail[(nb, 1 - S)] = [pack_send(f, D, S) for f in b[1]]
    for r in range(R):
        for S in range(2):
            b = bcs[r][D][S]
            if b[0] == "exchange":
                for f, buf in zip(b[1], mail[(r, S)]):
                    unpack_recv(f, D, S, buf)


def bc_world(grids: Sequence[Grid], bcs: Sequence, topos: Sequence[Topology] = None):
    """bc!(arch, grid, batchset) on every rank: D = N..1, side 1 then 2 (batch.jl:20-29)."""
    for D in reversed(range(grids[0].nd)):
        bc_dim_world(D, grids, bcs, topos)


def launch(launcher: Launcher, grid: Grid, op, args, bc=None):
    """single-rank convenience wrapper."""
    launch_world([launcher], [grid], op, [args], None if bc is None else [bc], None)


def bc_(grid: Grid, *field_bcs, exchange=None):
    """bc!(arch, grid, f => bc...; exchange) -- batch.jl:157 (single rank)."""
    bc_world([grid], [batch(grid, *field_bcs, exchange=exchange)], None)


# ------------------------------------------------------------------------------------------------ generic operators

def partial(grid, f, dim, *I):
    I = list(I) + [0] * (3 - len(I))
    return float(grid.B.lib.og_partial(C.byref(grid.c), C.byref(f.c), dim, *I))


def partial2(grid, f, dim, *I):
    I = list(I) + [0] * (3 - len(I))
    return float(grid.B.lib.og_partial2(C.byref(grid.c), C.byref(f.c), dim, *I))


def lerp(grid, f, to, *I):
    to = expand_loc(grid.nd, to)
    I = list(I) + [0] * (3 - len(I))
    return float(grid.B.lib.og_lerp(C.byref(grid.c), C.byref(f.c), (C.c_int32 * 3)(*(list(to) + [0] * (3 - grid.nd))), *I))


def dkd(grid, f, kf, dim, *I):
    I = list(I) + [0] * (3 - len(I))
    return float(grid.B.lib.og_dkd(C.byref(grid.c), C.byref(f.c), C.byref(kf.c), dim, *I))


def hlerp(grid, f, to, *I):
    to = expand_loc(grid.nd, to)
    I = list(I) + [0] * (3 - len(I))
    return float(grid.B.lib.og_hlerp(C.byref(grid.c), C.byref(f.c), (C.c_int32 * 3)(*(list(to) + [0] * (3 - grid.nd))), *I))


# field-level operators: dst[I] = OP(src...)[I] over [lo, hi] (default: the launch range [0, n+1]^N, KernelLaunch.jl:41,109)
OPER = {"left": 1, "right": 2, "delta": 3, "partial": 4, "partial2": 5, "dkd": 6, "lerp": 7, "hlerp": 8, "divg": 9,
        "lapl": 10, "divg_grad": 11, "vmag": 12, "grad": 13, "kgrad": 14}


def apply_operator(grid, kind, dst, src, k=None, dim=0, lo=None, hi=None):
    """dst / src: a Field or a sequence of Fields (vector components in x, y, z order)."""
    dst = [dst] if isinstance(dst, Field) else list(dst)
    src = [src] if isinstance(src, Field) else list(src)
    lo = [0] * grid.nd if lo is None else lo
    hi = [n + 1 for n in grid.n] if hi is None else hi
    blo, bhi = _box(grid.nd, lo, hi)
    grid.B.lib.og_apply_operator(C.byref(grid.c), OPER[kind] if isinstance(kind, str) else int(kind), int(dim), _fparr(dst),
                            _fparr(src), None if k is None else C.byref(k.c), blo, bhi)
