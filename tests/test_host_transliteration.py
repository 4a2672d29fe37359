"""
CPU-only: pins the index algebra of the "parity unpinned" rows (halo views, Launcher regions, distributed sub-grid) on
LITERAL transliterations of the reference's source text -- the reference asserts none of this anywhere, so the text
itself is the only anchor.  Each block below re-types the Julia expressions verbatim (1-based indices, Julia ranges)
and compares with BOTH independent restatements in this repo: the oracle (oracle/) and the product's host mirror
(chmy.jl_b200/).
"""
import itertools

import numpy as np
import pytest


class _NoArch:
    pass


# ------------------------------------------------------------------ src/Distributed/communication_views.jl:1-34
def jl_range(a, b):
    """Julia a:b (1-based, inclusive) -> numpy slice on a 0-based array"""
    return slice(a - 1, b)


def get_recv_view(side, D, array, halo_width):
    if side == 1:
        recv_range = jl_range(halo_width + 1, 2 * halo_width)                                        # :13
    else:
        recv_range = jl_range(array.shape[D] - 2 * halo_width + 1, array.shape[D] - halo_width)      # :19
    return tuple(recv_range if I == D else slice(None) for I in range(array.ndim))                   # Colon() elsewhere


def get_send_view(side, D, array, halo_width, overlap):
    if side == 1:
        send_range = jl_range(overlap + 2 * halo_width + 1, overlap + 3 * halo_width)                # :25
    else:
        send_range = jl_range(array.shape[D] - overlap - 3 * halo_width + 1, array.shape[D] - overlap - 2 * halo_width)  # :31
    return tuple(send_range if I == D else slice(None) for I in range(array.ndim))


@pytest.mark.parametrize("n,loc", [((9, 6), (1, 0)), ((9, 6), (0, 1)), ((7, 5, 4), (1, 0, 1)), ((7, 5, 4), (0, 0, 0))])
def test_halo_views_literally(oracle, n, loc):
    o = oracle
    g = o.Grid((0.0,) * len(n), (1.0,) * len(n), n)
    f = o.Field(g, loc)
    f.data[...] = np.arange(f.data.size, dtype=np.float64).reshape(f.sdims, order="F") + 1e3     # index-encoded parent array
    for D in range(len(n)):
        overlap = 1 if loc[D] == o.VERTEX else 0                                                  # :1-2
        for side in (1, 2):
            sv = get_send_view(side, D, f.data, 1, overlap)
            # exchange_halo.jl:41 copyto!(send_buf, send_view): column-major linear order of the view
            assert np.array_equal(o.pack_send(f, D, side - 1), f.data[sv].reshape(-1, order="F")), ("send", D, side)
            h = o.Field(g, loc)
            h.data[...] = f.data
            msg = -np.arange(1.0, f.data[sv].size + 1)
            o.unpack_recv(h, D, side - 1, msg)                                                    # :51 copyto!(recv_view, recv_buf)
            want = f.data.copy()
            rv = get_recv_view(side, D, want, 1)
            want[rv] = msg.reshape(want[rv].shape, order="F")
            assert np.array_equal(h.data, want), ("recv", D, side)


# ------------------------------------------------------------------ src/KernelLaunch.jl:56-87
def outer_worksize(worksize, outer_width, D):                                                    # :63-74, 1-based D and I
    return tuple(worksize[I - 1] if I < D else outer_width[I - 1] if I == D else worksize[I - 1] - 2 * outer_width[I - 1]
                 for I in range(1, len(worksize) + 1))


def outer_offset(worksize, outer_width, D, S):                                                   # :76-87
    return tuple(0 if I < D else ((0 if S == 1 else worksize[I - 1] - outer_width[I - 1]) if I == D else outer_width[I - 1])
                 for I in range(1, len(worksize) + 1))


@pytest.mark.parametrize("n,ow", [((256, 256), (16, 8)), ((16383, 16383), (128, 8)), ((767, 767, 767), (128, 8, 4)), ((30, 22, 14), (4, 3, 3))])
def test_launcher_regions_literally(oracle, n, ow):
    import chmy_b200 as ch
    o = oracle
    N = len(n)
    worksize = tuple(x + 2 for x in n)                                # :41 worksize = size(grid, Center()) .+ 2
    g = ch.UniformGrid(_NoArch(), origin=(0.0,) * N, extent=(1.0,) * N, dims=n)
    hl = ch.Launcher(_NoArch(), g, outer_width=ow)
    assert ch.worksize(hl) == worksize and ch.outer_width(hl) == ow
    assert ch.inner_worksize(hl) == tuple(w - 2 * q for w, q in zip(worksize, ow))               # :60
    assert ch.inner_offset(hl) == ow                                                             # :61
    regs = {name: (lo, hi) for name, lo, hi in o.Launcher(o.Grid((0.0,) * N, (1.0,) * N, n), ow).regions()}
    cover = np.zeros(worksize, dtype=np.int32) if np.prod(worksize) < 5e7 else None
    for D in range(1, N + 1):
        for S in (1, 2):
            size, off = outer_worksize(worksize, ow, D), outer_offset(worksize, ow, D, S)
            assert ch.outer_worksize(hl, D) == size and ch.outer_offset(hl, D, S) == off
            # launch: I = J + offset + Offset(-1), J in 1..size  ->  logical [off, off + size - 1]  (:109,163,172)
            lo, hi = regs[f"outer{D - 1}{S - 1}"]
            assert lo == off and hi == tuple(a + b - 1 for a, b in zip(off, size))
            if cover is not None:
                cover[tuple(slice(a, a + b) for a, b in zip(off, size))] += 1
    lo, hi = regs["inner"]
    assert lo == ow and hi == tuple(a + w - 2 * a - 1 for a, w in zip(ow, worksize))
    if cover is not None:                                             # the regions tile the launch range exactly once
        cover[tuple(slice(a, b + 1) for a, b in zip(lo, hi))] += 1
        assert (cover == 1).all()


# ------------------------------------------------------------------ src/Distributed/distributed_grid.jl:1-36, topology.jl:26-41
class _Topo:
    def __init__(self, dims, coords):
        self.dims, self.cart_coords = dims, coords

    def has_neighbor(self, D, S):                                     # non-periodic Cartesian grid (topology.jl:31)
        c = self.cart_coords[D - 1] + (-1 if S == 1 else 1)
        return 0 <= c < self.dims[D - 1]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("gdims,pdims", [((767 * 2, 767 * 2, 767 * 2), (2, 2, 2)), ((100, 37), (4, 2)), ((21, 20, 19), (2, 3, 1))])
def test_distributed_subgrid_literally(oracle, gdims, pdims, dtype):
    import chmy_b200 as ch
    from chmy_b200.architectures import DistributedArchitecture
    from chmy_b200.utils import fma_t
    o, N, T = oracle, len(gdims), np.dtype(dtype).type
    origin, extent = tuple(T(-1.0 - 0.25 * d) for d in range(N)), tuple(T(2.0 + 0.5 * d) for d in range(N))
    for rank, coords in enumerate(itertools.product(*[range(p) for p in pdims])):        # row-major ranks (MPI_Cart_create)
        topo = _Topo(pdims, coords)
        arch = DistributedArchitecture.__new__(DistributedArchitecture)
        arch.topology = topo
        g = ch.UniformGrid(arch, origin=origin, extent=extent, dims=gdims, dtype=dtype)
        ot = o.Topology(int(np.prod(pdims)), tuple(pdims), rank)
        assert ot.coords == coords
        og = o.local_grid(origin, extent, gdims, ot, dtype=dtype)
        for D in range(N):
            # UniformAxis(origin, extent, len) of the global axis (uniform_axis.jl:7-11)
            spacing = extent[D] / T(gdims[D])
            local_dims = -(-gdims[D] // pdims[D])                                         # cld (:25)
            offset = coords[D] * local_dims                                               # :26
            new_origin = T(fma_t(T(offset + 1 - 1), spacing, origin[D], dtype))           # vertex(ax, offset + 1) (:2; uniform_axis.jl:18)
            new_extent = spacing * T(local_dims)                                          # :3
            new_spacing = new_extent / T(local_dims)                                      # UniformAxis(new_origin, new_extent, len) recomputes it
            ax = g.axes[D]
            assert (ax.length, ax.origin, ax.extent, ax.spacing, ax.inv_spacing) == (local_dims, new_origin, new_extent, new_spacing, T(1.0) / new_spacing)
            assert (og.n[D], og.origin[D], og.extent[D], og.spacing[D], og.inv_spacing[D]) == \
                (local_dims, float(new_origin), float(new_extent), float(new_spacing), float(T(1.0) / new_spacing))
            for S in (1, 2):                                                              # overwrite_connectivity (:12-17)
                want = ch.Connected if topo.has_neighbor(D + 1, S) else ch.Bounded
                assert isinstance(ch.connectivity(g, D + 1, S), want)
                assert og.conn[D][S - 1] == (o.CONNECTED if topo.has_neighbor(D + 1, S) else o.BOUNDED)
                assert (ot.neighbors[D][S - 1] >= 0) == topo.has_neighbor(D + 1, S)
