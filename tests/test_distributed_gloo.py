"""CPU-only, world_size 2 and 4 over gloo: the host-side N>1 path -- bootstrap communicator, CartesianTopology
(coords / neighbours / shared rank), distributed sub-grid and the BatchSets it produces -- against the oracle's
independent restatement.  (The NCCL data path itself is covered by the -m gpu multi-GPU tests.)"""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ndims, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        import chmy_b200 as ch
        import oracle as o
        from chmy_b200.boundary_conditions import EmptyBatch, ExchangeBatch, FieldBatch, batch
        comm = ch.TorchDistComm()
        assert comm.bcast_bytes(bytes(range(128)) if rank == 0 else b"", 0) == bytes(range(128))
        topo = ch.CartesianTopology(comm, (0,) * ndims)
        ot = o.Topology(world, o.dims_create(world, (0,) * ndims), rank)
        assert topo.dims == ot.dims and topo.cart_coords == ot.coords and list(topo.neighbors) == ot.neighbors
        assert topo.shared_rank == rank and ch.node_size(topo) == world and ch.global_size(topo) == world
        for D in range(1, ndims + 1):
            for S in (1, 2):
                assert ch.has_neighbor(topo, D, S) == (ot.neighbors[D - 1][S - 1] >= 0)

        class A(ch.DistributedArchitecture):          # topology without a device: grid arithmetic is host-only
            def __init__(self, t):
                self.topology = t
        n_l = (30, 22, 14)[:ndims]
        n_g = tuple(a * p for a, p in zip(n_l, topo.dims))
        org, ext = (-1.0, -0.7, 0.3)[:ndims], (2.0, 1.9, 0.77)[:ndims]
        g = ch.UniformGrid(A(topo), origin=org, extent=ext, dims=n_g)
        og = o.local_grid(org, ext, n_g, ot)
        assert g.size(ch.Center()) == og.n == n_l
        d = g.desc()
        for a in range(ndims):
            assert d.origin[a] == og.origin[a] and d.extent[a] == og.extent[a]
            assert d.spacing[a] == og.spacing[a] and d.inv_spacing[a] == og.inv_spacing[a]
            assert [d.connectivity[a][s] for s in range(2)] == og.conn[a]

        class F(ch.Field):
            def __init__(self, name):
                self.name, self._h = name, None
        f = F("C")
        bs = batch(g, (f, ch.Neumann()), exchange=f)
        for a in range(ndims):
            for s in range(2):
                want = ExchangeBatch if og.conn[a][s] == o.CONNECTED else FieldBatch
                assert isinstance(bs[a][s], want)
        # gather!(dst, src, comm; root) -- gather.jl:9-33: index-encoded local blocks must land at coords .* size(src)
        import numpy as np
        loc_shape = (5, 4, 3)[:ndims]
        glob_shape = tuple(a * p for a, p in zip(loc_shape, topo.dims))
        G = np.arange(int(np.prod(glob_shape)), dtype=np.float64).reshape(glob_shape, order="F")
        mine = G[tuple(slice(c * n, (c + 1) * n) for c, n in zip(topo.cart_coords, loc_shape))].copy(order="F")
        for root in (0, world - 1):
            dst = np.full(glob_shape, -1.0, order="F") if rank == root else None
            ch.gather_(dst, mine, topo, root=root)
            if rank == root:
                assert np.array_equal(dst, G)
        if rank == 0:
            try:
                ch.gather_(np.zeros((1,) * ndims), mine, topo)
                raise AssertionError("size mismatch not detected")
            except ValueError:
                pass
        else:
            ch.gather_(None, mine, topo)
        q.put((rank, "ok"))
    except Exception as e:       # noqa
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ndims", [(2, 3), (2, 2), (4, 3)])
def test_topology_and_subgrid_over_gloo(world, ndims):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
