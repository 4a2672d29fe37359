"""
Distributed: Cartesian topology, halo exchange, scalar reductions.
Mirrors src/Distributed/{topology.jl:6-127, exchange_halo.jl:13-108} and the drivers' `max_mpi`
(examples/stokes_3d_inc_ve_T_mpi_perf.jl:16-19).  MPI is replaced by: an out-of-band communicator used ONLY to
bootstrap (rank, size, one 128-byte broadcast of the NCCL unique id), then NCCL over NVLink inside the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import socket

from . import _lib as L
from .fields import Field, FieldTuple

PROC_NULL = -1


class TorchDistComm:
    """Bootstrap communicator backed by torch.distributed (gloo or nccl), standing where MPI.COMM_WORLD stands in
    the reference's drivers.  Plumbing only: rank, size, broadcast of a few bytes."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self._group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def bcast_bytes(self, data: bytes, root: int = 0) -> bytes:
        box = [data if self.rank == root else None]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return box[0]

    def allgather_obj(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj, group=self._group)
        return out


def dims_create(nprocs: int, dims):
    """MPI.Dims_create(nprocs, dims) (topology.jl:28) through the C ABI (pure host code, no GPU needed)."""
    arr = (C.c_int32 * len(dims))(*[int(d) for d in dims])
    L.check(L.lib().chmy_dims_create(int(nprocs), len(dims), arr))
    return tuple(int(x) for x in arr)


class CartesianTopology:
    """CartesianTopology(comm, dims) -- topology.jl:26-41 (non-periodic, row-major ranks)."""

    def __init__(self, comm, dims):
        self.comm = comm
        self.nprocs = comm.size
        self.dims = dims_create(comm.size, dims)
        self.global_rank = comm.rank
        N = len(self.dims)
        c, r = [], comm.rank
        for d in reversed(self.dims):                      # MPI.Cart_coords
            c.append(r % d)
            r //= d
        self.cart_coords = tuple(reversed(c))
        nbs = []
        for D in range(N):                                 # MPI.Cart_shift(comm, D-1, 1) -> (left, right)
            pair = []
            for delta in (-1, 1):
                cc = list(self.cart_coords)
                cc[D] += delta
                if 0 <= cc[D] < self.dims[D]:
                    rr = 0
                    for ci, d in zip(cc, self.dims):
                        rr = rr * d + ci
                    pair.append(rr)
                else:
                    pair.append(PROC_NULL)
            nbs.append(tuple(pair))
        self.neighbors = tuple(nbs)
        self.node_name = socket.gethostname()              # MPI.Get_processor_name
        names = comm.allgather_obj(self.node_name) if hasattr(comm, "allgather_obj") else [self.node_name] * comm.size
        same = [i for i, n in enumerate(names) if n == self.node_name]
        self.shared_rank = same.index(comm.rank)           # MPI.Comm_split_type(SHARED) rank
        self.node_size_ = len(same)
        self._arch = None

    def _attach(self, child_arch):
        """Create the NCCL communicator inside the context (chmy_topo_create)."""
        uid = (C.c_uint8 * L.UNIQUE_ID_BYTES)()
        if self.nprocs > 1:
            if self.comm.rank == 0:
                L.check(L.lib().chmy_comm_unique_id(uid))
            raw = self.comm.bcast_bytes(bytes(uid), 0)
            uid = (C.c_uint8 * L.UNIQUE_ID_BYTES)(*raw)
        dims = (C.c_int32 * len(self.dims))(*self.dims)
        L.check(L.lib().chmy_topo_create(child_arch.ctx, self.nprocs, self.global_rank, len(self.dims), dims, uid))
        cc = (C.c_int32 * 3)()
        L.check(L.lib().chmy_topo_coords(child_arch.ctx, cc))
        nb = ((C.c_int32 * 2) * 3)()
        L.check(L.lib().chmy_topo_neighbors(child_arch.ctx, nb))
        assert tuple(cc[: len(self.dims)]) == self.cart_coords
        assert tuple(tuple(nb[D]) for D in range(len(self.dims))) == self.neighbors
        self._arch = child_arch

    def has_neighbor(self, dim: int, side: int) -> bool:
        """has_neighbor(topo, dim, side), 1-based (topology.jl:113)."""
        return self.neighbors[dim - 1][side - 1] != PROC_NULL

    def neighbor(self, dim: int, side: int) -> int:
        return self.neighbors[dim - 1][side - 1]


def global_rank(t):
    return t.global_rank


def shared_rank(t):
    return t.shared_rank


def node_name(t):
    return t.node_name


def dims(t):
    return t.dims


def cart_coords(t):
    return t.cart_coords


def neighbors(t):
    return t.neighbors


def neighbor(t, dim, side):
    return t.neighbor(dim, side)


def has_neighbor(t, dim, side):
    return t.has_neighbor(dim, side)


def global_size(t):
    return t.nprocs


def cart_comm(t):
    """cart_comm(topo) (topology.jl:66): the bootstrap communicator in Cartesian rank order (ranks are row-major over
    `dims`, exactly MPI.Cart_create's numbering without reordering); the data path uses the NCCL communicator."""
    return t.comm


def shared_comm(t):
    """shared_comm(topo) (topology.jl:73): the ranks of this node, as (node-local rank, size) of the bootstrap communicator."""
    return (t.shared_rank, t.node_size_)


def node_size(t):
    return t.node_size_


def _handles(fields):
    flat = []
    for f in fields:
        flat.extend(list(f) if isinstance(f, FieldTuple) else [f])
    arr = (C.c_void_p * len(flat))(*[f.handle for f in flat])
    return arr, len(flat)


def exchange_halo_(*args, blocking: bool = True):
    """exchange_halo!(side, dim, arch, grid, fields...)  -- exchange_halo.jl:13-61 (1-based side, dim)
    exchange_halo!(arch, grid, fields...)             -- exchange_halo.jl:73-84"""
    flags = L.LAUNCH_BLOCKING if blocking else L.LAUNCH_ASYNC
    if isinstance(args[0], int):
        side, dim, arch, grid, *fields = args
        arr, n = _handles(fields)
        g = grid.desc()
        L.check(L.lib().chmy_exchange_halo(arch.ctx, C.byref(g), int(dim) - 1, int(side) - 1, n, arr, flags))
    else:
        arch, grid, *fields = args
        arr, n = _handles(fields)
        g = grid.desc()
        L.check(L.lib().chmy_exchange_halo_all(arch.ctx, C.byref(g), n, arr, flags))


def allreduce_max(arch, *values):
    """MPI.Allreduce(x, MPI.MAX, comm) for a handful of scalars in ONE collective."""
    buf = (C.c_double * len(values))(*[float(v) for v in values])
    L.check(L.lib().chmy_allreduce_max(arch.ctx, buf, len(values)))
    return tuple(float(x) for x in buf)


def barrier(arch):
    L.check(L.lib().chmy_barrier(arch.ctx))


def gather_(*args, root: int = 0):
    """gather!(dst, src, comm; root=0) / gather!(arch, dst, src::Field; root) -- src/Distributed/gather.jl:9-42.

    Every rank contributes its local array (for a Field: the interior, copied to the host); on `root` block
    `cart_coords` of the Cartesian process grid lands at offset `cart_coords .* size(src)` of the column-major global
    array `dst`, whose size must be `size(src) .* dims` (the MPI subarray/Gatherv arithmetic of :13-33).  Non-root ranks
    pass `dst=None`.  Setup / output path of the reference (plots, dumps), not the PT loop: the blocks travel over the
    bootstrap communicator."""
    import numpy as np
    from .fields import interior
    if len(args) == 3 and hasattr(args[0], "topology"):
        arch, dst, src = args
        topo, local = arch.topology, np.asarray(interior(src))
    else:
        dst, src, topo = args
        local = np.asarray(src)
    blocks = topo.comm.allgather_obj((tuple(topo.cart_coords), local)) if topo.nprocs > 1 else [(tuple(topo.cart_coords), local)]
    if topo.global_rank != root:
        return
    if dst is None:
        raise ValueError("gather!: the root needs a destination array")
    want = tuple(n * p for n, p in zip(local.shape, topo.dims))
    if tuple(dst.shape) != want:
        raise ValueError(f"gather!: size(dst) = {tuple(dst.shape)} but size(src) .* dims = {want}")
    for coords, blk in blocks:
        if blk.shape != local.shape:
            raise ValueError("gather!: local arrays of different sizes")
        sl = tuple(slice(c * n, (c + 1) * n) for c, n in zip(coords, blk.shape))
        dst[sl] = blk
