"""
UniformGrid: host-side numbers only (sizes, coordinates, spacings, connectivity).
Mirrors src/Grids/{Grids.jl:26-57, uniform_axis.jl:1-37, abstract_axis.jl:10-36, structured_grid.jl:6-269} and the
distributed constructor src/Distributed/distributed_grid.jl:1-36.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .utils import fma_t


class Location:
    pass


class _Center(Location):
    code = L.CENTER

    def __repr__(self):
        return "Center()"


class _Vertex(Location):
    code = L.VERTEX

    def __repr__(self):
        return "Vertex()"


_C, _V = _Center(), _Vertex()


def Center():
    return _C


def Vertex():
    return _V


def flip(loc: Location) -> Location:
    """Grids.jl:40-43."""
    return _V if loc is _C else _C


class Bounded:
    code = L.BOUNDED

    def __repr__(self):
        return "Bounded()"


class Connected:
    code = L.CONNECTED

    def __repr__(self):
        return "Connected()"


def expand_loc(nd: int, loc):
    if isinstance(loc, Location):
        return (loc,) * nd
    loc = tuple(loc)
    if len(loc) != nd:
        raise ValueError(f"need {nd} locations, got {len(loc)}")
    return loc


class UniformAxis:
    """UniformAxis{T} (uniform_axis.jl:1-12).  Every number is a numpy scalar of the element type T (Float64 by default,
    Float32 as in the reference's tests) so that the host arithmetic rounds exactly where Julia's does."""

    def __init__(self, origin: float, extent: float, length: int, dtype=np.float64):
        T = self.T = np.dtype(dtype).type
        if T not in (np.float64, np.float32):
            raise TypeError("eltype(grid) must be Float64 or Float32 on this path")
        self.origin = T(origin)
        self.extent = T(extent)
        self.length = int(length)
        self.spacing = self.extent / T(self.length)       # :8
        self.inv_spacing = T(1.0) / self.spacing          # :9  inv(spacing)

    def nvertices(self):
        return self.length + 1

    def size(self, loc):
        return self.length + 1 if loc is _V else self.length

    def vertex(self, i: int) -> float:                    # :18
        return self.T(fma_t(self.T(i - 1), self.spacing, self.origin, self.T))

    def center(self, i: int) -> float:                    # :19
        return self.T(fma_t(self.T(i - 1), self.spacing, fma_t(self.T(0.5), self.spacing, self.origin, self.T), self.T))

    def coord(self, loc, i: int) -> float:
        return self.vertex(i) if loc is _V else self.center(i)

    def origin_at(self, loc):                             # :21-22
        return self.origin if loc is _V else self.T(fma_t(self.T(0.5), self.spacing, self.origin, self.T))

    def extent_at(self, loc):                             # :24-25
        return self.extent if loc is _V else self.extent - self.spacing


class StructuredGrid:
    """StructuredGrid{N,T,C,A} with uniform axes (structured_grid.jl:6-15)."""

    def __init__(self, axes, connectivity):
        self.axes = tuple(axes)
        self.connectivity_ = tuple(tuple(c) for c in connectivity)
        if len({ax.T for ax in self.axes}) != 1:
            raise TypeError("all axes of a grid share one element type")

    def eltype(self):
        """eltype(grid) (structured_grid.jl:6): numpy float64 | float32."""
        return self.axes[0].T

    # --- reference accessors
    def ndims(self):
        return len(self.axes)

    def size(self, loc):
        loc = expand_loc(self.ndims(), loc)
        return tuple(ax.size(l) for ax, l in zip(self.axes, loc))

    def desc(self) -> L.GridDesc:
        """The POD copied into every launch descriptor."""
        g = L.GridDesc()
        g.ndims = self.ndims()
        for d, ax in enumerate(self.axes):
            g.n[d] = ax.length
            # Float32 numbers widen exactly; the library rounds them back to the fields' element type
            g.origin[d], g.extent[d], g.spacing[d], g.inv_spacing[d] = (float(ax.origin), float(ax.extent), float(ax.spacing),
                                                                          float(ax.inv_spacing))
            for s in range(2):
                g.connectivity[d][s] = self.connectivity_[d][s].code
        return g


def UniformGrid(arch, *, origin, extent, dims, topology=None, dtype=np.float64) -> StructuredGrid:
    """UniformGrid(arch; origin, extent, dims, topology) -- structured_grid.jl:27-39.  On a distributed
    architecture `dims` is the GLOBAL size and the local sub-grid is returned (distributed_grid.jl:19-36).
    `dtype` stands for the type of the origin/extent numbers the Julia caller passes (`T(-5)`, test_grid_operators.jl:11)."""
    from .architectures import DistributedArchitecture
    N = len(dims)
    if not (len(origin) == len(extent) == N):
        raise ValueError("origin, extent and dims must have the same length")
    if topology is None:
        topology = tuple((Bounded(), Bounded()) for _ in range(N))
    axes = [UniformAxis(o, e, int(n), dtype) for o, e, n in zip(origin, extent, dims)]
    if not isinstance(arch, DistributedArchitecture):
        return StructuredGrid(axes, topology)
    topo = arch.topology
    local_dims = [-(-ax.length // p) for ax, p in zip(axes, topo.dims)]               # cld  (:25)
    offsets = [c * l for c, l in zip(topo.cart_coords, local_dims)]                    # :26
    local_axes = []
    for ax, off, ln in zip(axes, offsets, local_dims):                                 # subaxis :1-5
        local_axes.append(UniformAxis(ax.vertex(off + 1), ax.spacing * ax.T(ln), ln, ax.T))
    conn = tuple(tuple(Connected() if topo.has_neighbor(D + 1, S + 1) else topology[D][S] for S in range(2))
                 for D in range(N))                                                    # overwrite_connectivity :12-17
    return StructuredGrid(local_axes, conn)


def connectivity(grid, dim: int, side: int):
    """connectivity(grid, Dim(dim), Side(side)), 1-based (structured_grid.jl:57)."""
    return grid.connectivity_[dim - 1][side - 1]


def spacing(grid, *args):
    """spacing(grid) -> tuple (structured_grid.jl:161-162); spacing(grid, loc, Dim(d), i) -> scalar."""
    if not args:
        return tuple(ax.spacing for ax in grid.axes)
    return grid.axes[int(args[1]) - 1].spacing


def inv_spacing(grid, *args):
    if not args:
        return tuple(ax.inv_spacing for ax in grid.axes)
    return grid.axes[int(args[1]) - 1].inv_spacing


def coord(grid, loc, dim: int, i: int) -> float:
    """coord(grid, loc, Dim(dim), i) (structured_grid.jl:118-119)."""
    loc = expand_loc(grid.ndims(), loc)
    return grid.axes[dim - 1].coord(loc[dim - 1], i)


def coords(grid, loc, dim: int) -> np.ndarray:
    loc = expand_loc(grid.ndims(), loc)
    ax = grid.axes[dim - 1]
    return np.array([ax.coord(loc[dim - 1], i) for i in range(1, ax.size(loc[dim - 1]) + 1)], dtype=ax.T)


def centers(grid, dim=None):
    if dim is None:
        return tuple(coords(grid, Center(), d + 1) for d in range(grid.ndims()))
    return coords(grid, Center(), dim)


def vertices(grid, dim=None):
    if dim is None:
        return tuple(coords(grid, Vertex(), d + 1) for d in range(grid.ndims()))
    return coords(grid, Vertex(), dim)


def origin(grid, loc, dim: int):
    loc = expand_loc(grid.ndims(), loc)
    return grid.axes[dim - 1].origin_at(loc[dim - 1])


def extent(grid, loc, dim: int):
    loc = expand_loc(grid.ndims(), loc)
    return grid.axes[dim - 1].extent_at(loc[dim - 1])


def bounds(grid, loc, dim: int):
    o = origin(grid, loc, dim)
    return (o, o + extent(grid, loc, dim))


def axes_names(grid):
    return ("x", "y", "z")[:grid.ndims()]


# --- the remaining host-side accessors the reference exports (src/Grids/Grids.jl:8-12); pure numbers
def nvertices(grid_or_axis, dim: int | None = None):
    """nvertices(ax) = length + 1 (uniform_axis.jl:16; abstract_axis.jl:10)."""
    return grid_or_axis.nvertices() if dim is None else grid_or_axis.axes[dim - 1].nvertices()


def ncenters(grid_or_axis, dim: int | None = None):
    """ncenters(ax) = nvertices(ax) - 1 (abstract_axis.jl:11)."""
    return grid_or_axis.length if dim is None else grid_or_axis.axes[dim - 1].length


def axis(grid, dim: int) -> UniformAxis:
    """axis(grid, Dim(dim)) (structured_grid.jl:93)."""
    return grid.axes[dim - 1]


def vertex(grid, dim: int, i: int):
    """vertex(grid, Dim(dim), i) (structured_grid.jl:124)."""
    return grid.axes[dim - 1].vertex(i)


def center(grid, dim: int, i: int):
    """center(grid, Dim(dim), i) (structured_grid.jl:127)."""
    return grid.axes[dim - 1].center(i)


def direction(grid, name: str) -> int:
    """direction(grid, Val(:x)) = Dim(1) ... (structured_grid.jl:222-224); 1-based."""
    return ("x", "y", "z").index(name) + 1


def volume(grid, loc, *I):
    """volume(grid, loc, I...) = prod of the spacings (structured_grid.jl:248-268); folds left in eltype(grid)."""
    v = grid.axes[0].spacing
    for ax in grid.axes[1:]:
        v = v * ax.spacing
    return v


def inv_volume(grid, loc, *I):
    v = grid.axes[0].inv_spacing
    for ax in grid.axes[1:]:
        v = v * ax.inv_spacing
    return v
