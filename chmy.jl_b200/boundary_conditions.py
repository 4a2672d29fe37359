"""
Boundary-condition batches.  Mirrors src/BoundaryConditions/{first_order_boundary_condition.jl:9-40,
batch.jl:6-157}: `batch` normalises per-field specs into a BatchSet (host-side, tiny, rebuilt every iteration in
the drivers); the BatchSet is flattened into chmy_batch_desc[3][2] and applied by the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

from . import _lib as L
from .fields import Field, FieldTuple
from .grids import Connected, StructuredGrid
from .utils import AXES


class BoundaryFunction:
    """BoundaryFunction(fun; discrete=false, parameters=nothing, reduce_dims=true) -- boundary_function.jl:6-44,69-72.

    continuous: bf(grid, loc, dim, I...) = fun(reduce(dim, coord(grid, loc, I...))..., params...)
    discrete  : bf(grid, loc, dim, I...) = fun(grid, loc, dim, reduce(dim, I)..., params...)
    `loc` is ONE location (the field's location along the boundary dimension, batch.jl:174) applied to every axis,
    exactly as `coord(grid, loc::Location, I...)` does (structured_grid.jl:103-106).  `dim` is 1-based.
    A closure cannot cross the C ABI: `batch`/`bc!` evaluate it on the host into an (N-1)-dimensional device Field
    (value_field of chmy_batch_desc).  The reference evaluates the function inside every bc! kernel, so a closure over
    mutable or time-dependent state is re-evaluated here on every `batch` / `bc!` too (only the device buffer is re-used);
    `static=True` declares the function pure and lets the evaluated values be kept."""

    def __init__(self, fun, *, discrete: bool = False, parameters=None, reduce_dims: bool = True, static: bool = False):
        self.fun, self.discrete, self.parameters, self.reduce_dims = fun, bool(discrete), parameters, bool(reduce_dims)
        self.static = bool(static)
        self._cache = {}          # key -> (weakref to the architecture, value Field)

    def _params(self):
        if self.parameters is None:
            return ()
        return tuple(self.parameters) if isinstance(self.parameters, (tuple, list)) else (self.parameters,)

    def __call__(self, grid, loc, dim: int, *I):
        from .utils import remove_dim
        if len(I) != grid.ndims():
            raise ValueError("a boundary function takes one index per grid dimension")
        if self.discrete:
            J = remove_dim(dim, tuple(I)) if self.reduce_dims else tuple(I)
            return self.fun(grid, loc, dim, *J, *self._params())
        x = tuple(ax.coord(loc, i) for ax, i in zip(grid.axes, I))
        x = remove_dim(dim, x) if self.reduce_dims else x
        return self.fun(*x, *self._params())


@dataclass(frozen=True)
class FirstOrderBC:
    kind: int
    value: object = None      # None is the reference's `nothing` -> zero(eltype(grid)); Number; lower-dimensional
                              # Field (first_order_boundary_condition.jl:38-40); BoundaryFunction (boundary_function.jl)


def _bc_value(value):
    if value is None or isinstance(value, (Field, BoundaryFunction)):
        return value
    if callable(value):       # FirstOrderBC{Kind}(f::Function): continuous, reduced dims (boundary_function.jl:75-78)
        return BoundaryFunction(value)
    return float(value)


def Dirichlet(value=None) -> FirstOrderBC:
    return FirstOrderBC(L.DIRICHLET, _bc_value(value))


def Neumann(value=None) -> FirstOrderBC:
    return FirstOrderBC(L.NEUMANN, _bc_value(value))


def boundary_value_field(arch, grid: StructuredGrid, f: Field, bc: FirstOrderBC, D: int, S: int) -> Field:
    """Evaluate a BoundaryFunction at every face point the BC kernel visits (transverse indices 0..n_t+2,
    batch.jl:159-184) with the location / index the rule would pass (first_order_boundary_condition.jl:42-84):
    Dirichlet on a Vertex field: (Vertex, boundary node 1|d); Dirichlet on a Center field: (Center, halo 0|d+1);
    Neumann: (flip(loc), halo 0|d+1).  D, S are 0-based here."""
    from .grids import Center, Vertex, flip
    from .utils import insert_dim
    bf: BoundaryFunction = bc.value
    loc_f = f.loc[D]
    d = f.dims[D]
    if bc.kind == L.DIRICHLET and loc_f is Vertex():
        loc, idx = Vertex(), (1 if S == 0 else d)
    elif bc.kind == L.DIRICHLET:
        loc, idx = Center(), (0 if S == 0 else d + 1)
    else:
        loc, idx = flip(loc_f), (0 if S == 0 else d + 1)
    import weakref
    key = (id(arch), f.dtype.name, tuple((ax.origin, ax.extent, ax.length) for ax in grid.axes), D, S, bc.kind, loc.code, idx)
    hit = bf._cache.get(key)
    if hit is not None and hit[0]() is not arch:             # id() of a collected architecture re-used by another one
        hit = None
    if hit is not None and bf.static:
        return hit[1]
    N = grid.ndims()
    taxes = [ax for a, ax in enumerate(grid.axes) if a != D]
    tgrid = StructuredGrid(taxes, [c for a, c in enumerate(grid.connectivity_) if a != D])
    # d_t = n_t + 1: logical indices -1..n_t+3 exist; the device buffer of an earlier evaluation is re-used
    vf = hit[1] if hit is not None else Field(arch, tgrid, Vertex(), f.dtype)
    ext = [ax.length + 3 for ax in taxes]                     # indices 0..n_t+2
    import numpy as np
    vals = np.empty(ext, dtype=f.dtype, order="F")
    for J in np.ndindex(*ext):
        I = insert_dim(D + 1, tuple(int(j) for j in J), idx)
        vals[J] = bf(grid, loc, D + 1, *I)
    vf.from_host(vals, [0] * (N - 1), [e - 1 for e in ext])
    bf._cache[key] = (weakref.ref(arch), vf)
    return vf


class EmptyBatch:
    def __repr__(self):
        return "EmptyBatch()"


class FieldBatch:
    def __init__(self, fields, conditions):
        self.fields, self.conditions = tuple(fields), tuple(conditions)

    def __repr__(self):
        return f"FieldBatch ({len(self.fields)} fields)"


class ExchangeBatch:
    def __init__(self, fields):
        self.fields = tuple(fields)

    def __repr__(self):
        return f"ExchangeBatch ({len(self.fields)} fields)"


def _regularise_impl(N, bc):
    """batch.jl:126-130."""
    if isinstance(bc, FirstOrderBC):
        return [(bc, bc)] * N
    out = [(None, None)] * N
    for name, v in dict(bc).items():
        D = AXES.index(name)
        out[D] = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    return out


def _regularise_exchange(N, exchange):
    """batch.jl:135-141."""
    if exchange is None:
        return [None] * N
    if isinstance(exchange, Field):
        return [(exchange,)] * N
    if isinstance(exchange, FieldTuple):
        # a NamedTuple keyed by axis names: dim D exchanges only component D (batch.jl:139-141)
        out = [None] * N
        for name in exchange.keys():
            out[AXES.index(name)] = (getattr(exchange, name),)
        return out
    if isinstance(exchange, dict):
        out = [None] * N
        for name, f in exchange.items():
            out[AXES.index(name)] = (f,) if isinstance(f, Field) else tuple(f)
        return out
    return [tuple(exchange)] * N


def batch(grid: StructuredGrid, *field_bcs, exchange=None):
    """batch(grid, f => bc...; exchange) -> BatchSet = ((left, right) per dim) -- batch.jl:72-116.
    `f => bc` is written as the pair (f, bc); a per-axis spec is a dict {'x': bc | (l, r), ...}."""
    N = grid.ndims()
    fields = [fb[0] for fb in field_bcs]
    bcs = [_regularise_impl(N, fb[1]) for fb in field_bcs]
    exch = _regularise_exchange(N, exchange)
    out = []
    for D in range(N):
        sides = []
        for S in range(2):
            if isinstance(grid.connectivity_[D][S], Connected):              # batch_impl(::Connected ...) :98-101
                e = exch[D]
                sides.append(ExchangeBatch(e) if e and any(x is not None for x in e) else EmptyBatch())
            else:                                                            # batch_impl(::Bounded ...)  :103-105
                fb = [(f, b[D][S]) for f, b in zip(fields, bcs) if b[D][S] is not None]      # prune :151-155
                sides.append(FieldBatch(*zip(*fb)) if fb else EmptyBatch())
        out.append(tuple(sides))
    return tuple(out)


def fill_batch_desc(dst: L.BatchDesc, b, arch=None, grid=None, D=None, S=None):
    if isinstance(b, FieldBatch):
        if len(b.fields) > L.MAX_BATCH_FIELDS:
            raise ValueError(f"a FieldBatch holds at most {L.MAX_BATCH_FIELDS} fields on this path")
        dst.kind, dst.nfields = L.BATCH_FIELD, len(b.fields)
        for q, (f, bc) in enumerate(zip(b.fields, b.conditions)):
            dst.fields[q] = f.handle
            dst.bc_kind[q] = bc.kind
            v = bc.value
            if isinstance(v, BoundaryFunction):
                if arch is None or grid is None:
                    raise ValueError("a BoundaryFunction-valued condition needs the architecture and the grid to be lowered")
                v = boundary_value_field(arch, grid, f, bc, D, S)
            if isinstance(v, Field):
                if v.ndims() != f.ndims() - 1:
                    raise ValueError("a Field-valued condition takes a field with one dimension less than the grid")
                dst.value[q], dst.value_field[q] = 0.0, v.handle
            else:
                dst.value[q], dst.value_field[q] = (0.0 if v is None else v), None
    elif isinstance(b, ExchangeBatch):
        if len(b.fields) > L.MAX_BATCH_FIELDS:
            raise ValueError(f"an ExchangeBatch holds at most {L.MAX_BATCH_FIELDS} fields on this path")
        dst.kind, dst.nfields = L.BATCH_EXCHANGE, len(b.fields)
        for q, f in enumerate(b.fields):
            dst.fields[q] = f.handle
    else:
        dst.kind, dst.nfields = L.BATCH_EMPTY, 0


def batchset_array(batchset, arch=None, grid=None):
    arr = ((L.BatchDesc * 2) * L.MAX_DIMS)()
    for D, sides in enumerate(batchset):
        for S in range(2):
            fill_batch_desc(arr[D][S], sides[S], arch, grid, D, S)
    return arr


def bc_(arch, grid: StructuredGrid, *field_bcs, exchange=None, blocking: bool = True):
    """bc!(arch, grid, f => bc...; exchange) and bc!(arch, grid, batchset) -- batch.jl:20-29,157."""
    if len(field_bcs) == 1 and isinstance(field_bcs[0], tuple) and field_bcs[0] and isinstance(field_bcs[0][0], tuple) \
            and not isinstance(field_bcs[0][0][0], Field):
        bs = field_bcs[0]                                                    # already a BatchSet
    else:
        bs = batch(grid, *field_bcs, exchange=exchange)
    g = grid.desc()
    arr = batchset_array(bs, arch, grid)
    L.check(L.lib().chmy_bc(arch.ctx, C.byref(g), arr, L.LAUNCH_BLOCKING if blocking else L.LAUNCH_ASYNC))
