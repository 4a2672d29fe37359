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


@dataclass(frozen=True)
class FirstOrderBC:
    kind: int
    value: Optional[float] = None      # None is the reference's `nothing` -> zero(eltype(grid))


def Dirichlet(value=None) -> FirstOrderBC:
    return FirstOrderBC(L.DIRICHLET, None if value is None else float(value))


def Neumann(value=None) -> FirstOrderBC:
    return FirstOrderBC(L.NEUMANN, None if value is None else float(value))


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


def fill_batch_desc(dst: L.BatchDesc, b):
    if isinstance(b, FieldBatch):
        if len(b.fields) > L.MAX_BATCH_FIELDS:
            raise ValueError(f"a FieldBatch holds at most {L.MAX_BATCH_FIELDS} fields on this path")
        dst.kind, dst.nfields = L.BATCH_FIELD, len(b.fields)
        for q, (f, bc) in enumerate(zip(b.fields, b.conditions)):
            dst.fields[q] = f.handle
            dst.bc_kind[q] = bc.kind
            dst.value[q] = 0.0 if bc.value is None else bc.value
    elif isinstance(b, ExchangeBatch):
        if len(b.fields) > L.MAX_BATCH_FIELDS:
            raise ValueError(f"an ExchangeBatch holds at most {L.MAX_BATCH_FIELDS} fields on this path")
        dst.kind, dst.nfields = L.BATCH_EXCHANGE, len(b.fields)
        for q, f in enumerate(b.fields):
            dst.fields[q] = f.handle
    else:
        dst.kind, dst.nfields = L.BATCH_EMPTY, 0


def batchset_array(batchset):
    arr = ((L.BatchDesc * 2) * L.MAX_DIMS)()
    for D, sides in enumerate(batchset):
        for S in range(2):
            fill_batch_desc(arr[D][S], sides[S])
    return arr


def bc_(arch, grid: StructuredGrid, *field_bcs, exchange=None, blocking: bool = True):
    """bc!(arch, grid, f => bc...; exchange) and bc!(arch, grid, batchset) -- batch.jl:20-29,157."""
    if len(field_bcs) == 1 and isinstance(field_bcs[0], tuple) and field_bcs[0] and isinstance(field_bcs[0][0], tuple) \
            and not isinstance(field_bcs[0][0][0], Field):
        bs = field_bcs[0]                                                    # already a BatchSet
    else:
        bs = batch(grid, *field_bcs, exchange=exchange)
    g = grid.desc()
    arr = batchset_array(bs)
    L.check(L.lib().chmy_bc(arch.ctx, C.byref(g), arr, L.LAUNCH_BLOCKING if blocking else L.LAUNCH_ASYNC))
