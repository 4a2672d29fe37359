"""
Fields: padded device storage owned by the C ABI + the host-visible accessors of the reference.
Mirrors src/Fields/{field.jl:6-209, function_field.jl:12-65}.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .grids import Center, Location, StructuredGrid, Vertex, expand_loc


class AbstractField:
    pass


class Field(AbstractField):
    """Field(backend|arch, grid, loc; halo=1) -- field.jl:56-74.  Storage is allocated (zero-filled) by
    chmy_field_create; logical indexing follows field.jl:18-22."""

    def __init__(self, arch, grid: StructuredGrid, loc=None, dtype=None, *, halo: int = 1, layout: int = L.LAYOUT_PITCHED):
        if halo != 1:
            raise NotImplementedError("this path implements the default halo=1 of the reference")
        loc = Center() if loc is None else loc
        self.arch = arch
        self.grid = grid
        # Field(backend, grid, loc, type=eltype(grid)) -- field.jl:56-57
        self.dtype = np.dtype(grid.eltype() if dtype is None else dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("eltype(field) must be Float64 or Float32 on this path")
        self.loc = expand_loc(grid.ndims(), loc)
        self.dims = grid.size(self.loc)
        nd = len(self.dims)
        h = C.c_void_p()
        L.check(L.lib().chmy_field_create_typed(arch.ctx, nd, L.i64x3(self.dims, 1), L.i32x3([l.code for l in self.loc]),
                                                layout, L.F32 if self.dtype == np.float32 else L.F64, C.byref(h)))
        self._h = h

    @classmethod
    def shell(cls, grid: StructuredGrid, loc=None, dtype=None, *, layout: int = L.LAYOUT_PITCHED) -> "Field":
        """Descriptor-only Field (chmy_field_create_shell): location, sizes, layout and element type without an
        architecture or device storage.  Good for `Launcher.validate` -- checking `op => args` against the library's
        argument rules on a machine without a GPU -- and refused by everything that touches storage."""
        self = cls.__new__(cls)
        loc = Center() if loc is None else loc
        self.arch, self.grid = None, grid
        self.dtype = np.dtype(grid.eltype() if dtype is None else dtype)
        self.loc = expand_loc(grid.ndims(), loc)
        self.dims = grid.size(self.loc)
        h = C.c_void_p()
        L.check(L.lib().chmy_field_create_shell(len(self.dims), L.i64x3(self.dims, 1), L.i32x3([l.code for l in self.loc]),
                                                layout, L.F32 if self.dtype == np.float32 else L.F64, C.byref(h)))
        self._h = h
        return self

    # ------------------------------------------------------------------ handles
    @property
    def handle(self):
        if self._h is None:
            raise L.ChmyError("field was freed")
        return self._h

    def info(self) -> L.FieldInfo:
        fi = L.FieldInfo()
        L.check(L.lib().chmy_field_get_info(self.handle, C.byref(fi)))
        return fi

    def free(self):
        if getattr(self, "_h", None) is not None:
            L.lib().chmy_field_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ------------------------------------------------------------------ AbstractArray-ish interface
    def size(self):
        return self.dims

    def ndims(self):
        return len(self.dims)

    def _box(self, pad: int):
        lo = [1 - pad] * self.ndims()
        hi = [d + pad for d in self.dims]
        return lo, hi

    def to_host(self, lo, hi, out: np.ndarray | None = None) -> np.ndarray:
        shape = tuple(h - l + 1 for l, h in zip(lo, hi))
        if out is None:
            out = np.empty(shape, dtype=self.dtype, order="F")
        elif out.shape != shape or out.dtype != self.dtype or not out.flags.f_contiguous:
            raise ValueError(f"to_host(out=...): need a Fortran-contiguous {self.dtype} array of shape {shape}")
        L.check(L.lib().chmy_field_copy_to_host(self.arch.ctx, self.handle, out.ctypes.data_as(C.c_void_p),
                                                L.i64x3(lo), L.i64x3(hi)))
        return out

    def from_host(self, arr: np.ndarray, lo, hi):
        shape = tuple(h - l + 1 for l, h in zip(lo, hi))
        a = np.asfortranarray(np.broadcast_to(np.asarray(arr, dtype=self.dtype), shape))
        L.check(L.lib().chmy_field_copy_from_host(self.arch.ctx, self.handle, a.ctypes.data_as(C.c_void_p),
                                                  L.i64x3(lo), L.i64x3(hi)))

    def parent(self) -> np.ndarray:
        """Array(parent(f)): the whole padded array dims+4 (field.jl:16)."""
        return self.to_host(*self._box(2))

    def __getitem__(self, I):
        I = (I,) if isinstance(I, int) else tuple(I)
        return float(self.to_host(I, I).reshape(-1)[0])


def halo(f: Field) -> int:
    return 1


def location(f, dim=None):
    return f.loc if dim is None else f.loc[dim - 1]


def interior(f: Field, with_halo: bool = False, out: np.ndarray | None = None) -> np.ndarray:
    """Array(interior(f; with_halo)) -- field.jl:33-37 (host copy; the reference returns a device view).
    `out`: copy into this (e.g. pinned, see `pinned_array`) host array instead of allocating one."""
    return f.to_host(*f._box(1 if with_halo else 0), out=out)


class _PinnedOwner:
    """Keeps a chmy_host_alloc buffer alive for as long as a numpy view of it exists."""

    def __init__(self, arch, ptr):
        self.arch, self.ptr = arch, ptr

    def __del__(self):
        try:
            if self.ptr and getattr(self.arch, "_ctx", True) is not None:
                L.lib().chmy_host_free(self.arch.ctx, self.ptr)
        except Exception:
            pass
        self.ptr = None


def pinned_array(arch, shape, dtype=np.float64) -> np.ndarray:
    """Fortran-ordered host array in page-locked memory (chmy_host_alloc): the fast host side of
    set!(f, A) / Array(interior(f)).  Freed when the last numpy view of it is garbage-collected."""
    shape = tuple(int(s) for s in shape)
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    p = C.c_void_p()
    L.check(L.lib().chmy_host_alloc(arch.ctx, n * dtype.itemsize, C.byref(p)))
    owner = _PinnedOwner(arch, p)
    buf = (C.c_byte * (n * dtype.itemsize)).from_address(p.value)
    buf._chmy_owner = owner                    # the ctypes object is the numpy array's base: ties the lifetimes
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order="F")


def parent(f: Field) -> np.ndarray:
    return f.parent()


def fill_parent_(f: Field, v: float):
    """fill!(parent(f), v) as used by test/test_fields.jl:22."""
    lo, hi = f._box(2)
    L.check(L.lib().chmy_field_fill(f.arch.ctx, f.handle, float(v), L.i64x3(lo), L.i64x3(hi)))


class _InitIncl:
    """The `init_incl` closure of the Stokes drivers (examples/stokes_3d_inc_ve_T.jl:125):
    ifelse(sum((x - x0)^2) < r^2, in, out).  It is the one function body evaluated on the device."""

    def __call__(self, *a):
        nd = (len(a) - 3) // 2
        xs, c0, (r, inn, out) = a[:nd], a[nd:2 * nd], a[2 * nd:]
        s = None
        for x, c in zip(xs, c0):
            t = (x - c) * (x - c)
            s = t if s is None else s + t
        return np.where(s < r * r, inn, out)


init_incl = _InitIncl()


class _InitGauss:
    """The Gaussian initial condition of the diffusion drivers, `(x, y) -> exp(-x^2 - y^2)` (examples/diffusion_2d_mpi.jl:46,
    diffusion_2d_mpi_perf.jl:56), in 1-3 dimensions.  `set!(C, grid, init_gauss)` evaluates it on the device."""

    def __call__(self, *xs):
        s = None
        for x in xs:
            s = -(x * x) if s is None else s - x * x
        return np.exp(s)


init_gauss = _InitGauss()


def _incl_struct(nd, loc, parameters) -> L.Inclusion:
    p = dict(parameters)
    names = ["x0", "y0", "z0"][:nd]
    s = L.Inclusion()
    s.active = 1
    for d in range(nd):
        s.loc[d] = loc[d].code
        s.c0[d] = float(p[names[d]])
    s.r, s.inn, s.out = float(p["r"]), float(p["in"]), float(p["out"])
    return s


def set_(f, *args, discrete: bool = False, parameters=()):
    """set!(f, val) / set!(f, A) / set!(f, other) / set!(f, grid, fun; parameters) -- field.jl:87-145."""
    if isinstance(f, FieldTuple):
        for c in f:
            set_(c, *args, discrete=discrete, parameters=parameters)
        return
    lo, hi = f._box(0)
    if len(args) == 1:
        a = args[0]
        if isinstance(a, Field):                                             # field.jl:109-119
            if a.dims != f.dims:
                raise ValueError("set!(f, other): size mismatch")
            L.check(L.lib().chmy_field_copy(f.arch.ctx, f.handle, a.handle, L.i64x3(lo), L.i64x3(hi)))
        elif np.isscalar(a):                                                 # field.jl:87
            L.check(L.lib().chmy_field_fill(f.arch.ctx, f.handle, float(a), L.i64x3(lo), L.i64x3(hi)))
        else:                                                                # field.jl:98
            a = np.asarray(a, dtype=f.dtype)
            if a.shape != tuple(f.dims):
                raise ValueError(f"set!(f, A): A has shape {a.shape}, interior is {f.dims}")
            f.from_host(a, lo, hi)
        return
    grid, fun = args
    params = tuple(parameters.values()) if isinstance(parameters, dict) else tuple(parameters)
    if discrete:                                                             # field.jl:126-129: fun(grid, loc, I..., params...)
        vals = np.empty(tuple(f.dims), dtype=f.dtype, order="F")
        for I in np.ndindex(*vals.shape):
            vals[I] = fun(grid, f.loc, *[int(i) + 1 for i in I], *params)
        f.from_host(vals, lo, hi)
        return
    if fun is init_gauss and not params:                                     # device kernel (CUDA exp: ~1e-16 from the host's)
        g = grid.desc()
        L.check(L.lib().chmy_field_set_gaussian(f.arch.ctx, f.handle, C.byref(g)))
        return
    if fun is init_incl:                                                     # device kernel, bit-identical coords
        inc = _incl_struct(grid.ndims(), f.loc, parameters)
        g = grid.desc()
        L.check(L.lib().chmy_field_set_inclusion(f.arch.ctx, f.handle, C.byref(g), C.byref(inc)))
        return
    # any other host callable: evaluate on the host at the exact (muladd) coordinates and upload the bits
    from .grids import coords
    cs = [coords(grid, f.loc, d + 1) for d in range(grid.ndims())]
    mesh = np.meshgrid(*cs, indexing="ij")
    f.from_host(np.asarray(fun(*mesh, *params), dtype=f.dtype), lo, hi)


class FieldTuple:
    """NamedTuple of Fields (VectorField / TensorField): attribute access + ordered iteration."""

    def __init__(self, **fields):
        self._names = tuple(fields)
        self.__dict__.update(fields)

    def __iter__(self):
        return (getattr(self, n) for n in self._names)

    def __len__(self):
        return len(self._names)

    def keys(self):
        return self._names

    def __getitem__(self, k):
        return getattr(self, self._names[k] if isinstance(k, int) else k)


def vector_location(dim: int, N: int):
    """field.jl:148."""
    return tuple(Vertex() if i == dim else Center() for i in range(1, N + 1))


def VectorField(arch, grid: StructuredGrid, **kw) -> FieldTuple:
    """field.jl:161-169."""
    N = grid.ndims()
    return FieldTuple(**{"xyz"[D - 1]: Field(arch, grid, vector_location(D, N), **kw) for D in range(1, N + 1)})


def TensorField(arch, grid: StructuredGrid, **kw) -> FieldTuple:
    """field.jl:182-206."""
    Cn, Vx = Center(), Vertex()
    if grid.ndims() == 2:
        return FieldTuple(xx=Field(arch, grid, Cn, **kw), yy=Field(arch, grid, Cn, **kw), xy=Field(arch, grid, Vx, **kw))
    if grid.ndims() == 3:
        return FieldTuple(xx=Field(arch, grid, Cn, **kw), yy=Field(arch, grid, Cn, **kw), zz=Field(arch, grid, Cn, **kw),
                          xy=Field(arch, grid, (Vx, Vx, Cn), **kw), xz=Field(arch, grid, (Vx, Cn, Vx), **kw),
                          yz=Field(arch, grid, (Cn, Vx, Vx), **kw))
    raise ValueError("TensorField is defined for 2D and 3D grids")


class FunctionField(AbstractField):
    """FunctionField(func, grid, loc; discrete=false, parameters) -- function_field.jl:12-59: a field whose value at I is
    `func(coord(grid, loc, I)..., params...)` (continuous) or `func(grid, loc, I..., params...)` (discrete).

    A closure cannot cross the C ABI.  The drivers' `init_incl` (the one body the named solvers use, mpi_perf.jl:141-142)
    is evaluated IN-KERNEL from coordinates (chmy_inclusion) and needs no storage; any other function is evaluated on the
    host at the exact (muladd) coordinates of every index a kernel can read -- 0..d+1 per dim -- and uploaded once into a
    stored Field that takes the FunctionField's place in the launch (same values, one field of memory)."""

    def __init__(self, func, grid: StructuredGrid, loc, *, discrete: bool = False, parameters=None):
        self.func = func
        self.grid = grid
        self.loc = expand_loc(grid.ndims(), loc)
        self.discrete = bool(discrete)
        self.parameters = dict(parameters) if isinstance(parameters, dict) else parameters
        self.dims = grid.size(self.loc)
        self._stored = None

    def in_kernel(self) -> bool:
        return self.func is init_incl and not self.discrete

    def inclusion(self) -> L.Inclusion:
        return _incl_struct(self.grid.ndims(), self.loc, self.parameters or {})

    def _params(self):
        p = self.parameters
        if p is None:
            return ()
        return tuple(p.values()) if isinstance(p, dict) else (tuple(p) if isinstance(p, (tuple, list)) else (p,))

    def values(self) -> np.ndarray:
        """func at logical indices 0..d+1 of every dim (interior + halo), in eltype(grid)."""
        from .grids import coord
        nd, T = self.grid.ndims(), self.grid.eltype()
        idx = [range(0, d + 2) for d in self.dims]
        if self.discrete:                                                    # function_field.jl:53-59
            out = np.empty([d + 2 for d in self.dims], dtype=T, order="F")
            for I in np.ndindex(*out.shape):
                out[I] = self.func(self.grid, self.loc, *[int(i) for i in I], *self._params())
            return out
        cs = [np.array([self.grid.axes[d].coord(self.loc[d], i) for i in idx[d]], dtype=T) for d in range(nd)]
        mesh = np.meshgrid(*cs, indexing="ij")                               # function_field.jl:49-51
        return np.asfortranarray(np.broadcast_to(np.asarray(self.func(*mesh, *self._params()), dtype=T), mesh[0].shape))

    def materialize(self, arch) -> "Field":
        if self._stored is None or self._stored.arch is not arch:
            f = Field(arch, self.grid, self.loc)
            f.from_host(self.values(), [0] * len(self.dims), [d + 1 for d in self.dims])
            self._stored = f
        return self._stored


class ConstantField(AbstractField):
    """ConstantField{T} (src/Fields/constant_field.jl:1-41): a 0-dimensional field with the same value at every index.
    As the `rho_g` argument of update_velocity! it travels as a chmy_inclusion whose inside and outside values coincide
    (evaluated in-kernel, no storage)."""

    def __init__(self, value, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.value = self.dtype.type(value)

    def size(self):
        return ()

    def ndims(self):
        return 0

    def __getitem__(self, I):
        return self.value

    def inclusion_at(self, grid: StructuredGrid, loc) -> L.Inclusion:
        s = L.Inclusion()
        s.active = 1
        for d, l in enumerate(expand_loc(grid.ndims(), loc)):
            s.loc[d] = l.code
            s.c0[d] = 0.0
        s.r, s.inn, s.out = 0.0, float(self.value), float(self.value)      # sum((x - c0)^2) < 0 never holds -> `out`
        return s

    def __repr__(self):
        return f"{type(self).__name__}{{{self.dtype.name}}}({self.value})"


class ZeroField(ConstantField):
    def __init__(self, dtype=np.float64):
        super().__init__(0.0, dtype)


class OneField(ConstantField):
    def __init__(self, dtype=np.float64):
        super().__init__(1.0, dtype)


class ValueField(ConstantField):
    pass


def maxabs(f: Field, with_halo: bool = False) -> float:
    """maximum(abs.(interior(f))) fused into one reduction kernel (drivers: stokes_3d_inc_ve_T.jl:158,172-175)."""
    lo, hi = f._box(1 if with_halo else 0)
    out = C.c_double()
    L.check(L.lib().chmy_field_maxabs(f.arch.ctx, f.handle, L.i64x3(lo), L.i64x3(hi), C.byref(out)))
    return float(out.value)


def maxabs_many(*fields: Field, with_halo: bool = False):
    """The maxima of several fields in ONE device round trip (the residual check of the drivers,
    stokes_3d_inc_ve_T.jl:171-175: four `maximum(abs.(interior(f)))`, each with its own host synchronisation there)."""
    if not fields:
        return ()
    n = len(fields)
    arch = fields[0].arch
    lo, hi = (C.c_int64 * (3 * n))(), (C.c_int64 * (3 * n))()
    hs = (C.c_void_p * n)(*[f.handle for f in fields])
    for q, f in enumerate(fields):
        if f.arch is not arch:
            raise ValueError("maxabs_many: the fields must live on one architecture")
        l, h = f._box(1 if with_halo else 0)
        l3, h3 = L.i64x3(l), L.i64x3(h)
        for a in range(3):
            lo[3 * q + a], hi[3 * q + a] = l3[a], h3[a]
    out = (C.c_double * n)()
    L.check(L.lib().chmy_field_maxabs_many(arch.ctx, n, hs, lo, hi, out))
    return tuple(float(x) for x in out)
