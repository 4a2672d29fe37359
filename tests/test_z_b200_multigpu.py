"""
(Named test_z_* so that it runs after the single-GPU suites: each case spawns one process per GPU.)
Multi-GPU parity (NCCL over NVLink, one process per GPU) against the oracle's lock-step simulation of the same
Cartesian process grid.  Needs >= 2 CUDA devices; skipped otherwise.

The reference pins nothing on this path (no golden vector for exchange_halo!), so the checks
are: (1) index-encoded fields (rank*1e6 + linear storage index) through exchange_halo! must equal the oracle's
restatement of communication_views.jl:1-34 bit for bit, corners and padding included; (2) the decomposed Stokes /
diffusion solvers (split launch: inner region on the main stream, slabs + BC + exchange on the boundary stream)
must reproduce the oracle's world run per rank, full padded arrays, <= 1e-12 relative (they are bit-identical in
practice) and the residual history.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    sys.path.insert(0, ROOT)
    import ctypes as C
    import chmy_b200
    n = C.c_int(0)
    rc = chmy_b200.load_library().chmy_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _encode(shape, rank):
    return rank * 1.0e6 + np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape, order="F")


def _cmp(name, a, b, tol):
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if tol == 0.0:
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        if not same.all():
            i = tuple(np.argwhere(~same)[0])
            raise AssertionError(f"{name}: {int((~same).sum())} cells differ; first at {i}: oracle {a[i]!r} cuda {b[i]!r}")
        return 0.0
    err = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    assert err <= tol, f"{name}: relative error {err:.3e} > {tol:.1e}"
    return err


def _worker(rank, world, port, case, q):
    # every rank also runs the oracle's lock-step world on the host: share the cores instead of oversubscribing them
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import install_dryrun_if_requested
    install_dryrun_if_requested()       # tests/test_gpu_suite_dryrun.py only (CHMY_DRYRUN=1); never on the GPU box
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    arch = None
    try:
        import chmy_b200 as ch
        import oracle as o
        import drivers as OD
        from chmy_b200 import drivers as BD
        kind, n = case
        nd = len(n)
        peer = kind.endswith("+peer")        # EXPERIMENTAL transport: peer stores + sequence flags instead of NCCL send/recv
        kind = kind[:-5] if peer else kind
        arch = ch.Arch(ch.B200Backend(), ch.TorchDistComm(), (0,) * nd, device_id=rank + 1)
        if peer:
            ch.set_exchange_mode(arch, "peer")
        pd = arch.topology.dims
        assert pd == o.dims_create(world, (0,) * nd)
        if kind == "exchange":
            # examples/exchange_halo.jl:20-30 with index-encoded contents instead of the rank id
            topos = [o.Topology(world, pd, r) for r in range(world)]
            n_g = tuple(a * p for a, p in zip(n, pd))
            org, ext = (-1.0,) * nd, (2.0,) * nd
            ogs = [o.local_grid(org, ext, n_g, t) for t in topos]
            g = ch.UniformGrid(arch, origin=org, extent=ext, dims=n_g)
            locs = [(0,) * nd, (1,) + (0,) * (nd - 1), (0,) * (nd - 1) + (1,), (1,) * nd]
            ofs = [[o.Field(og, l) for l in locs] for og in ogs]
            for r in range(world):
                for f in ofs[r]:
                    f.data[...] = _encode(f.sdims, r)
            bfs = [ch.Field(arch, g, tuple(ch.Vertex() if x else ch.Center() for x in l)) for l in locs]
            for f, of in zip(bfs, ofs[rank]):
                f.from_host(of.data.copy(), [-1] * nd, [d + 2 for d in of.dims])
            o.bc_world(ogs, [o.batch(ogs[r], exchange=tuple(ofs[r])) for r in range(world)], topos)
            ch.exchange_halo_(arch, g, *bfs)
            for f, of, l in zip(bfs, ofs[rank], locs):
                _cmp(f"exchange loc={l}", of.data, f.parent(), 0.0)
            if peer:
                # again, several times, on changed contents: slots alternate, the flags keep counting, nothing stale survives
                for rep in range(5):
                    for r in range(world):
                        for f in ofs[r]:
                            f.data[...] = _encode(f.sdims, r + 10 * (rep + 1))
                    for f, of in zip(bfs, ofs[rank]):
                        f.from_host(of.data.copy(), [-1] * nd, [d + 2 for d in of.dims])
                    o.bc_world(ogs, [o.batch(ogs[r], exchange=tuple(ofs[r])) for r in range(world)], topos)
                    ch.exchange_halo_(arch, g, *bfs)
                    ch.synchronize(arch)
                    for f, of, l in zip(bfs, ofs[rank], locs):
                        _cmp(f"exchange rep={rep} loc={l}", of.data, f.parent(), 0.0)
                sent_peer, sent_nccl = ch.exchange_stats(arch)
                assert sent_peer > 0 and sent_nccl == 0, (sent_peer, sent_nccl)      # every link mapped: nothing fell back
            q.put((rank, "ok", 0.0))
            return
        if kind == "gather":
            # gather!(arch, dst, field) (src/Distributed/gather.jl:36-42): the interiors of every rank's field, block by block in
            # Cartesian order, on the root -- Center and Vertex locations (a Vertex field's shared boundary vertex appears twice,
            # as in the reference: size(dst) = size(interior) .* dims)
            n_g = tuple(a * p for a, p in zip(n, pd))
            g = ch.UniformGrid(arch, origin=(-1.0,) * nd, extent=(2.0,) * nd, dims=n_g)
            coords = arch.topology.cart_coords
            for loc in [(0,) * nd, (1,) + (0,) * (nd - 1), (1,) * nd]:
                f = ch.Field(arch, g, tuple(ch.Vertex() if x else ch.Center() for x in loc))
                shape = tuple(f.dims)
                mine = rank * 1.0e6 + np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape, order="F")
                ch.set_(f, mine)
                for root in (0, world - 1):
                    dst = np.full(tuple(a * p for a, p in zip(shape, pd)), np.nan, order="F") if rank == root else None
                    ch.gather_(arch, dst, f, root=root)
                    if rank == root:
                        for r in range(world):
                            c, rr = [], r
                            for dsz in reversed(pd):
                                c.append(rr % dsz); rr //= dsz
                            c = tuple(reversed(c))
                            want = r * 1.0e6 + np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape, order="F")
                            sl = tuple(slice(ci * m, (ci + 1) * m) for ci, m in zip(c, shape))
                            assert np.array_equal(dst[sl], want), (loc, root, r)
                if rank == 0:                                    # collective call; the root's destination has the wrong size
                    with pytest.raises(ValueError):
                        ch.gather_(arch, np.zeros((1,) * nd), f, root=0)
                else:
                    ch.gather_(arch, None, f, root=0)
            q.put((rank, "ok", 0.0))
            return
        if kind == "diffusion":
            ow = (16, 8)
            rngs = [np.random.default_rng(100 + r).random(n) for r in range(world)]
            osol = OD.Diffusion2D(n, proc_dims=pd, outer_width=ow, C0=rngs)
            bsol = BD.Diffusion2D(arch, n, outer_width=ow, C0=rngs[rank], blocking=False)
            osol.run(25)
            bsol.run(25)
            hist_ok = True
        else:
            ow = (8, 4, 3)[:nd]
            if kind == "stokes_fused":      # the lazily fused stress+velocity sweep under the split launch + exchange
                ch.set_fusion(arch, True)
            osol = OD.Stokes(n, proc_dims=pd, rho_g_function=True, outer_width=ow, adv_coef=0.01, re_m=2.5 * np.pi)
            bsol = BD.Stokes(arch, n, rho_g_function=True, outer_width=ow, adv_coef=0.01, re_m=2.5 * np.pi, blocking=False)
            ho = osol.run(2, 40, 10)
            hb = bsol.run(2, 40, 10)
            assert len(ho) == len(hb) == 8
            for a, b in zip(ho, hb):
                assert a[:2] == b[:2]
                for x, y in zip(a[2:], b[2:]):
                    assert abs(x - y) <= 1e-12 * abs(x), (a, b)
            assert bsol.dt == osol.dt and bsol.eta_ve == osol.eta_ve
            if kind == "stokes_fused":
                assert ch.fused_count(arch) == 80 + 40       # mechanics sweeps + the thermal sweeps of the second outer step
        if peer:
            sent_peer, sent_nccl = ch.exchange_stats(arch)
            assert sent_peer > 0 and sent_nccl == 0, (sent_peer, sent_nccl)
        worst = 0.0
        bf = bsol.fields()
        for k, f in osol.fields(rank).items():
            worst = max(worst, _cmp(f"rank {rank} {k}", f.data, bf[k].parent(), 1e-12))
        q.put((rank, "ok", worst))
    except Exception:       # noqa
        import traceback
        q.put((rank, traceback.format_exc(), None))
    finally:
        try:
            if arch is not None:
                arch.close()
        finally:
            dist.destroy_process_group()


CASES = [
    (2, ("exchange", (9, 7, 5))),
    (2, ("exchange", (12, 9))),
    (2, ("stokes", (30, 22, 14))),
    (2, ("stokes", (40, 33))),
    (2, ("stokes_fused", (30, 22, 14))),
    (4, ("stokes_fused", (24, 22, 14))),
    (8, ("stokes_fused", (24, 20, 16))),
    (2, ("diffusion", (64, 48))),
    (2, ("gather", (9, 7, 5))),
    (2, ("gather", (12, 9))),
    (4, ("gather", (9, 7, 5))),
    (4, ("exchange", (9, 7, 5))),
    (4, ("stokes", (24, 22, 14))),
    (8, ("exchange", (9, 7, 5))),
    (8, ("stokes", (24, 20, 16))),
]
# the fused sweep is the default of the drivers and the bench: it gets every world size
CASES.sort(key=lambda c: (c[0], c[1][0]))
# peer-store transport (comm.cu, CHMY_EXCHANGE_PEER; opt-in): protocol proven on the CPU (tests/test_peer_protocol.py)
PEER_CASES = [
    (2, ("exchange+peer", (9, 7, 5))),
    (2, ("exchange+peer", (12, 9))),
    (2, ("stokes_fused+peer", (30, 22, 14))),
    (2, ("diffusion+peer", (64, 48))),
    (4, ("exchange+peer", (9, 7, 5))),
    (4, ("stokes_fused+peer", (24, 22, 14))),
    (8, ("exchange+peer", (9, 7, 5))),
    (8, ("stokes_fused+peer", (24, 20, 16))),
]
CASES += PEER_CASES          # green on 2-, 4- and 8-GPU boxes since round 2 (profiles/r2_c9_*, r2_c11_*)


@pytest.mark.parametrize("world,case", CASES, ids=[f"{w}gpu-{c[0]}-{'x'.join(map(str, c[1]))}" for w, c in CASES])
def test_multigpu_matches_oracle_world(world, case):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=300) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for rank, msg, _ in res:
        assert msg == "ok", f"rank {rank}: {msg}"
