"""
TEST INFRASTRUCTURE -- a dry-run stand-in for libchmy_b200.so, used ONLY by tests/test_gpu_suite_dryrun.py.

There is no GPU where the CPU suite runs, yet most of what can break a `-m gpu` test file is not CUDA: the test's own
Python, the host mirror's flattening of fields / batches / descriptors, element-type plumbing, the drivers.  This module
restates the C ABI's entry points on top of the CPU oracle so that the single-GPU `-m gpu` test files can be EXECUTED
on the CPU: every `chmy_*` call the mirror makes lands here, descriptors are decoded by the field order documented in
include/chmy_b200.h (an independent reading of the ABI: a mirror that flattens `op => args` in the wrong order computes
the wrong thing here too), argument validation is done by the REAL library (chmy_validate_launch on descriptor-only twins
of the fields), and the arithmetic is the oracle's.  Numerically such a run compares the oracle with itself -- it proves
nothing about the CUDA kernels and is never reported as parity; it proves that the GPU suites are runnable programs.

Not a product path: nothing under chmy.jl_b200/ knows about it; it is injected by tests/conftest.py only when
CHMY_DRYRUN=1 is set (by the dry-run test's child process).
"""
import ctypes as C

import numpy as np

NT = {2: 3, 3: 6}


class _Field:
    def __init__(self, o, nd, dims, loc, layout, dtype, shell):
        self.nd, self.dims, self.loc, self.layout, self.dtype = nd, tuple(dims), tuple(loc), layout, np.dtype(dtype)
        n = tuple(d - l for d, l in zip(dims, loc))
        g = o.Grid((0.0,) * nd, (1.0,) * nd, n, dtype=dtype)
        self.f = o.Field(g, tuple(loc))
        self.shell = shell                       # handle of the descriptor-only twin in the real library
        self.has_storage = True


class DryRunLib:
    """Duck-typed replacement of the ctypes CDLL: same function names, ctypes arguments, integer status returns."""

    def __init__(self, real, oracle):
        self.real, self.o = real, oracle
        self.fields, self.ctxs, self.bufs = {}, {}, {}
        self.next = 0x1000
        self.err = b""
        self.launches = 0
        self.disable_fast = 0
        import os
        self.ndev = int(os.environ.get("CHMY_DRYRUN_NGPU", "1"))     # "devices" the multi-rank dry run pretends to have

    # ------------------------------------------------------------------ helpers
    def _id(self):
        self.next += 0x100
        return self.next

    @staticmethod
    def _set(ref, value):
        ref._obj.value = value

    @staticmethod
    def _h(x):
        return x.value if hasattr(x, "value") else x

    def _fail(self, code, msg):
        self.err = msg.encode()
        return code

    def _F(self, h):
        return self.fields[self._h(h)]

    def _real_error(self, rc):
        self.err = self.real.chmy_last_error()
        return rc

    def _grid(self, gd, dtype):
        nd = gd.ndims
        conn = [[gd.connectivity[d][s] for s in range(2)] for d in range(nd)]
        g = self.o.Grid([gd.origin[d] for d in range(nd)], [gd.extent[d] for d in range(nd)], [gd.n[d] for d in range(nd)],
                        conn, dtype=dtype)
        for d in range(nd):      # the descriptor carries the numbers the host computed: they must be the oracle's
            assert g.spacing[d] == gd.spacing[d] and g.inv_spacing[d] == gd.inv_spacing[d], "grid numbers differ from the oracle's"
        return g

    def _box(self, fld, lo, hi):
        nd = fld.nd
        for a in range(nd):
            if lo[a] < -1 or hi[a] > fld.dims[a] + 2:
                return None
        return tuple(slice(lo[a] + 1, hi[a] + 2) for a in range(nd))

    def _host(self, ptr, fld, sl_shape):
        n = int(np.prod(sl_shape))
        ct = C.c_float if fld.dtype == np.float32 else C.c_double
        addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
        return np.ctypeslib.as_array((ct * n).from_address(addr)).reshape(sl_shape, order="F")

    # ------------------------------------------------------------------ library / context
    def chmy_abi_version(self):
        return self.real.chmy_abi_version()

    def chmy_last_error(self):
        return self.err

    def chmy_struct_size(self, which):
        return self.real.chmy_struct_size(which)

    def chmy_device_count(self, out):
        self._set(out, self.ndev)
        return 0

    def chmy_ctx_create(self, dev, out):
        if not 1 <= dev <= self.ndev:
            return self._fail(-1, f"device_id {dev} out of range 1..{self.ndev}")
        i = self._id()
        self.ctxs[i] = {"fuse": 0}
        self._set(out, i)
        return 0

    def chmy_ctx_destroy(self, ctx):
        self.ctxs.pop(self._h(ctx), None)
        return 0

    def chmy_synchronize(self, ctx):
        self._flush_all()
        return 0

    def chmy_ctx_launch_count(self, ctx, out):
        self._set(out, self.launches)
        return 0

    def chmy_event_record(self, ctx, slot):
        self._flush_all()
        return 0

    def chmy_time_fused_sweep(self, ctx, a, b):
        return 0

    def chmy_event_elapsed_ms(self, ctx, a, b, out):
        self._set(out, 1.0)
        return 0

    def chmy_set_tuning(self, a, b):
        if b >= 0:
            self.true_div = int(b)
        return 0

    def chmy_set_launch_tuning(self, ctx, overlap, bc_fold):
        if overlap >= 0:
            self.ctxs[self._h(ctx)]["overlap"] = int(overlap)
        return 0

    def chmy_fusion_fallback_count(self, ctx, out):
        self._set(out, 0)
        return 0

    def chmy_overlapped_count(self, ctx, out):
        self._set(out, self.ctxs[self._h(ctx)].get("noverl", 0))
        return 0

    def chmy_set_exchange_mode(self, ctx, mode):          # transport only: results cannot depend on it
        self.ctxs[self._h(ctx)]["xmode"] = int(mode)
        return 0

    def chmy_exchange_stats(self, ctx, peer, nccl):
        c = self.ctxs[self._h(ctx)]
        self._set(peer, c.get("peer_msgs", 0))
        self._set(nccl, c.get("nccl_msgs", 0))
        return 0

    def chmy_set_fused_tuning(self, *a):
        return 0

    def chmy_set_fused2d_tuning(self, *a):
        return 0

    # The lazily fused launches (api.cu: chmy_launch): only the PROTOCOL is restated -- which launch is deferred, what
    # flushes it, which pair counts as one sweep -- because the fused suites assert the sweep counter.  The arithmetic of
    # a "fused" pair is simply the two oracle ops (the dry run never claims anything about the sweep kernels).
    PAIR = {4: 5, 1: 2, 6: 7}

    def chmy_set_fusion(self, ctx, enable):
        c = self.ctxs[self._h(ctx)]
        c["fuse"] = ((enable & 3) or 1) if enable else 0
        c["pending"] = None
        return 0

    def chmy_fused_count(self, ctx, out):
        self._set(out, self.ctxs[self._h(ctx)].get("nfused", 0))
        return 0

    def _flush_all(self):
        for c in self.ctxs.values():
            c["pending"] = None

    def _fusion_step(self, ctx, d, F):
        c = self.ctxs[self._h(ctx)]
        fuse, nd = c.get("fuse", 0), d.grid.ndims
        pend = c.get("pending")
        odd = bool((d.flags & 2) and d.has_bc and d.has_outer_width and
                   ((d.outer_width[0] & 1) or ((d.grid.n[0] + 2 - d.outer_width[0]) & 1)))
        pitched = all(f is None or f.layout == 0 for f in F)
        H = [d.fields[q] for q in range(d.nfields)]
        nt = NT.get(nd, 0)
        if pend is not None and self.PAIR.get(pend[0]) == d.op and pitched and not odd:
            P = pend[1]
            if d.op == 5:      # same tau, Pr, V as the deferred stress launch (chmy_fused_eligible)
                same = P[:nt] == H[2 * nd + 1:2 * nd + 1 + nt] and P[nt] == H[2 * nd] and P[nt + 2:nt + 2 + nd] == H[:nd]
            elif d.op == 2:    # compute_q: q.x q.y C ; update_C: C q.x q.y
                same = P[:2] == H[1:3] and P[2] == H[0]
            else:              # thermal_flux: qT[nd] T V[nd] ; thermal: T T_old qT[nd]
                same = P[:nd] == H[2:2 + nd] and P[nd] == H[0] and H[1] != H[0]
            if same:
                c["nfused"] = c.get("nfused", 0) + 1
                if d.op == 5 and nd == 3:      # ops_fused.cu: division mode of the sweep (velocity scalars: eta_ve nudtau; stress: eta eta_ve G dt)
                    S = pend[2]
                    e = C.c_int32(0)
                    two = all(self.real.chmy_division_two_op_exact(float(x), C.byref(e)) == 0 and e.value for x in (S[2] * S[3], S[0], S[1], 3.0))
                    c["div_mode"] = 1 if getattr(self, "true_div", 0) else (2 if two else 0)
                any_ex = any(d.bc[D][S].kind == 2 for D in range(nd) for S in range(2))
                if d.op == 5 and nd == 3 and d.has_bc and not (d.flags & 2) and (c.get("overlap", 1) == 2 or (c.get("overlap", 1) and any_ex)):
                    c["noverl"] = c.get("noverl", 0) + 1      # api.cu run_overlapped: the batches run behind the boundary tiles
        defer = (not d.has_bc) and pitched and (
            ((fuse & 1) and d.op == 4 and nd == 3) or
            ((fuse & 2) and ((nd == 2 and d.op in (4, 1, 6)) or (nd == 3 and d.op == 6))))
        c["pending"] = (d.op, H, [d.scalars[q] for q in range(d.nscalars)]) if defer else None

    def chmy_selftest_division2(self, ctx, c, n, seed, bad, proved):
        e = C.c_int32(0)
        self.real.chmy_division_two_op_exact(float(c), C.byref(e))      # the proof is a pure host function: the real one
        self._set(bad, 0)
        self._set(proved, int(e.value))
        return 0

    def chmy_last_division_mode(self, ctx, out):
        self._set(out, self.ctxs[self._h(ctx)].get("div_mode", 0))
        return 0

    def chmy_division_two_op_exact(self, c, out):
        return self.real.chmy_division_two_op_exact(c, out) if getattr(self, "real", None) is not None else (self._set(out, 1) or 0)

    def chmy_selftest_division(self, ctx, c, n, seed, bad, used):
        self._set(bad, 0)
        self._set(used, 1)
        return 0

    # ---- ranks: the bootstrap group the tests / bench create (torch.distributed, gloo) carries the dry run's "NCCL" traffic
    @staticmethod
    def _dist():
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() else None

    def chmy_comm_unique_id(self, out):
        return 0

    def chmy_topo_create(self, ctx, nranks, rank, nd, dims, uid):
        d = [dims[a] for a in range(nd)]
        t = self.o.Topology(nranks, tuple(d), rank)
        self.ctxs[self._h(ctx)]["topo"] = t
        return 0

    def chmy_topo_coords(self, ctx, out):
        t = self.ctxs[self._h(ctx)]["topo"]
        for a, c in enumerate(t.coords):
            out[a] = c
        return 0

    def chmy_topo_neighbors(self, ctx, out):
        t = self.ctxs[self._h(ctx)]["topo"]
        for a, (l, r) in enumerate(t.neighbors):
            out[a][0], out[a][1] = l, r
        return 0

    def chmy_allreduce_max(self, ctx, buf, n):
        self._flush_all()
        dist = self._dist()
        if dist is not None and dist.get_world_size() > 1:
            import torch
            t = torch.tensor([buf[i] for i in range(n)], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for i in range(n):
                buf[i] = float(t[i])
        return 0

    def chmy_barrier(self, ctx):
        self._flush_all()
        dist = self._dist()
        if dist is not None and dist.get_world_size() > 1:
            dist.barrier()
        return 0

    def _exchange_dim(self, ctx, D, sides):
        """exchange_halo.jl:13-61 for both sides of one dim: my side S talks to the neighbour's side 1-S"""
        import torch
        dist, topo = self._dist(), self.ctxs[self._h(ctx)].get("topo")
        if dist is None or topo is None:
            raise RuntimeError("halo exchange requested but the architecture has no topology")
        reqs, recv = [], {}
        for S, fields in sides.items():
            nb = topo.neighbors[D][S]
            if nb < 0:
                raise RuntimeError("no neighbor to communicate")
            msg = np.concatenate([np.asarray(self.o.pack_send(f, D, S), dtype=np.float64) for f in fields])
            c = self.ctxs[self._h(ctx)]           # one message per (dim, side), counted under the chosen transport
            key = "peer_msgs" if c.get("xmode", 0) == 1 else "nccl_msgs"
            c[key] = c.get(key, 0) + 1
            reqs.append(dist.isend(torch.from_numpy(msg.copy()), nb, tag=2 * D + (1 - S)))
            recv[S] = torch.empty(msg.size, dtype=torch.float64)
            reqs.append(dist.irecv(recv[S], nb, tag=2 * D + S))
        for r in reqs:
            r.wait()
        for S, fields in sides.items():
            off = 0
            for f in fields:
                n = int(np.prod([s for a, s in enumerate(f.sdims) if a != D]))
                self.o.unpack_recv(f, D, S, recv[S][off:off + n].numpy().astype(f.dtype))
                off += n

    def _apply_batches(self, ctx, g, bc):
        """bc!(arch, grid, batchset): D = N..1; FieldBatch sides first, then the exchange of that dim (batch.jl:20-29)"""
        for D in reversed(range(g.nd)):
            ex = {}
            for S in range(2):
                b = bc[D][S]
                if b.kind == 1:
                    fb = []
                    for q in range(b.nfields):
                        vf = b.value_field[q]
                        fb.append((self._F(b.fields[q]).f, self.o.BC(b.bc_kind[q], self._F(vf).f if vf else b.value[q])))
                    self.o.bc_side(g, D, S, ("field", fb))
                elif b.kind == 2:
                    ex[S] = [self._F(b.fields[q]).f for q in range(b.nfields)]
            if ex:
                self._exchange_dim(ctx, D, ex)

    def chmy_exchange_halo(self, ctx, gd, dim, side, nf, fields, flags):
        self._flush_all()
        self._exchange_dim(ctx, dim, {side: [self._F(fields[q]).f for q in range(nf)]})
        return 0

    def chmy_exchange_halo_all(self, ctx, gd, nf, fields, flags):
        self._flush_all()
        gd = gd._obj
        fs = [self._F(fields[q]).f for q in range(nf)]
        for D in reversed(range(gd.ndims)):
            ex = {S: fs for S in range(2) if gd.connectivity[D][S] == 1}
            if ex:
                self._exchange_dim(ctx, D, ex)
        return 0

    def chmy_dims_create(self, nranks, nd, dims):
        return self.real.chmy_dims_create(nranks, nd, dims)

    def chmy_host_alloc(self, ctx, nbytes, out):
        b = C.create_string_buffer(nbytes)
        self.bufs[C.addressof(b)] = b
        self._set(out, C.addressof(b))
        return 0

    def chmy_host_free(self, ctx, p):
        self.bufs.pop(self._h(p), None)
        return 0

    # ------------------------------------------------------------------ fields
    def _create(self, nd, dims, loc, layout, dtype, out, storage):
        shell = C.c_void_p()
        rc = self.real.chmy_field_create_shell(nd, dims, loc, layout, dtype, C.byref(shell))      # the real argument checks
        if rc:
            return self._real_error(rc)
        fld = _Field(self.o, nd, [dims[a] for a in range(nd)], [loc[a] for a in range(nd)], layout,
                     np.float32 if dtype == 1 else np.float64, shell)
        fld.has_storage = storage
        i = self._id()
        self.fields[i] = fld
        self._set(out, i)
        return 0

    def chmy_field_create_typed(self, ctx, nd, dims, loc, layout, dtype, out):
        self._flush_all()
        if self._h(ctx) not in self.ctxs:
            return self._fail(-1, "bad context")
        return self._create(nd, dims, loc, layout, dtype, out, True)

    def chmy_field_create_shell(self, nd, dims, loc, layout, dtype, out):
        return self._create(nd, dims, loc, layout, dtype, out, False)

    def chmy_field_destroy(self, h):
        self._flush_all()
        fld = self.fields.pop(self._h(h), None)
        if fld is not None:
            self.real.chmy_field_destroy(fld.shell)
        return 0

    def chmy_field_get_info(self, h, out):
        fld = self._F(h)
        rc = self.real.chmy_field_get_info(fld.shell, out)          # dims, strides, layout, dtype from the real layout code
        if rc == 0 and fld.has_storage:
            es = fld.dtype.itemsize
            info = out._obj
            lead = (128 // es - 1) if fld.layout == 0 else 0
            base = 0x7F0000000000 + (0 if fld.layout == 0 else 8)    # a plausible 128-B aligned allocation
            off = lead + 2 + (2 * info.stride[1] if fld.nd > 1 else 0) + (2 * info.stride[2] if fld.nd > 2 else 0)
            info.origin_ptr = base + es * off
            info.base_ptr = base + es * lead
        return rc

    def chmy_field_fill(self, ctx, h, v, lo, hi):
        self._flush_all()
        fld = self._F(h)
        sl = self._box(fld, lo, hi)
        if sl is None or not fld.has_storage:
            return self._fail(-1, "box outside the padded field")
        fld.f.data[sl] = v
        return 0

    def chmy_field_copy_from_host(self, ctx, h, src, lo, hi):
        self._flush_all()
        fld = self._F(h)
        sl = self._box(fld, lo, hi)
        if sl is None or not fld.has_storage:
            return self._fail(-1, "box outside the padded field")
        shape = fld.f.data[sl].shape
        if int(np.prod(shape)) > 0:
            fld.f.data[sl] = self._host(src, fld, shape)
        return 0

    def chmy_field_copy_to_host(self, ctx, h, dst, lo, hi):
        self._flush_all()
        fld = self._F(h)
        sl = self._box(fld, lo, hi)
        if sl is None or not fld.has_storage:
            return self._fail(-1, "box outside the padded field")
        shape = fld.f.data[sl].shape
        if int(np.prod(shape)) > 0:
            self._host(dst, fld, shape)[...] = fld.f.data[sl]
        return 0

    def chmy_field_copy(self, ctx, hd, hs, lo, hi):
        self._flush_all()
        d, s = self._F(hd), self._F(hs)
        if d.nd != s.nd:
            return self._fail(-1, "set!(f, other): dimensionality mismatch")
        if d.dtype != s.dtype:
            return self._fail(-1, "set!(f, other): element types differ")
        sl = self._box(d, lo, hi)
        if sl is None:
            return self._fail(-1, "box outside the padded field")
        d.f.data[sl] = s.f.data[sl]
        return 0

    def chmy_field_set_inclusion(self, ctx, h, gd, inc):
        self._flush_all()
        fld = self._F(h)
        g, q = self._grid(gd._obj, fld.dtype), inc._obj
        # evaluate on a field bound to the launch grid's numbers (the coordinates come from the grid)
        tmp = self.o.Field(g, fld.loc)
        tmp.data[...] = fld.f.data
        self.o.set_inclusion(tmp, self.o.Inclusion(fld.loc, tuple(q.c0[d] for d in range(fld.nd)), q.r, q.inn, q.out))
        fld.f.data[...] = tmp.data
        return 0

    def chmy_field_set_gaussian(self, ctx, h, gd):
        self._flush_all()
        fld = self._F(h)
        g = self._grid(gd._obj, fld.dtype)
        cs = [np.array([g.coord(d, fld.loc[d], i) for i in range(1, fld.f.dims[d] + 1)], dtype=fld.dtype) for d in range(fld.nd)]
        mesh = np.meshgrid(*cs, indexing="ij")
        s = None
        for x in mesh:
            s = -(x * x) if s is None else s - x * x
        inner = tuple(slice(2, -2) for _ in range(fld.nd))
        fld.f.data[inner] = np.exp(s)
        return 0

    def chmy_field_maxabs(self, ctx, h, lo, hi, out):
        self._flush_all()
        fld = self._F(h)
        sl = self._box(fld, lo, hi)
        if sl is None or not fld.has_storage:
            return self._fail(-1, "box outside the padded field")
        a = np.abs(fld.f.data[sl])
        self._set(out, float("nan") if np.isnan(a).any() else (float(a.max()) if a.size else 0.0))
        return 0

    def chmy_field_maxabs_many(self, ctx, n, hs, lo, hi, out):
        self._flush_all()
        if not 1 <= n <= 64:
            return self._fail(-1, "between 1 and 64 fields per call")
        for q in range(n):
            one = C.c_double()
            rc = self.chmy_field_maxabs(ctx, hs[q], (C.c_int64 * 3)(*lo[3 * q:3 * q + 3]), (C.c_int64 * 3)(*hi[3 * q:3 * q + 3]), C.byref(one))
            if rc:
                return rc
            out[q] = one.value
        return 0

    # ------------------------------------------------------------------ halo slabs (host helpers of the parity tests)
    def chmy_halo_slab_len(self, h, dim, out):
        self._set(out, int(np.prod([s for a, s in enumerate(self._F(h).f.sdims) if a != dim])))
        return 0

    def chmy_halo_pack(self, ctx, h, dim, side, buf):
        self._flush_all()
        fld = self._F(h)
        ref = self.o.pack_send(fld.f, dim, side)
        self._host(buf, fld, (ref.size,))[...] = ref
        return 0

    def chmy_halo_unpack(self, ctx, h, dim, side, buf):
        self._flush_all()
        fld = self._F(h)
        n = int(np.prod([s for a, s in enumerate(fld.f.sdims) if a != dim]))
        self.o.unpack_recv(fld.f, dim, side, np.array(self._host(buf, fld, (n,))))
        return 0

    # ------------------------------------------------------------------ batches / launch
    def _validate(self, d):
        """chmy_validate_launch of the REAL library on a copy of the descriptor whose handles are the shell twins"""
        from chmy_b200 import _lib as L
        t = L.LaunchDesc.from_buffer_copy(d)
        for q in range(d.nfields):
            t.fields[q] = self._F(d.fields[q]).shell if d.fields[q] else None
        for D in range(3):
            for S in range(2):
                b = t.bc[D][S]
                for q in range(b.nfields if b.kind else 0):
                    b.fields[q] = self._F(d.bc[D][S].fields[q]).shell
                    if d.bc[D][S].value_field[q]:
                        b.value_field[q] = self._F(d.bc[D][S].value_field[q]).shell
        rc = self.real.chmy_validate_launch(C.byref(t))
        return self._real_error(rc) if rc else 0

    def chmy_validate_launch(self, dref):
        return self._validate(dref._obj)

    def chmy_bc(self, ctx, gd, arr, flags):
        self._flush_all()
        gd = gd._obj
        dt = None
        for D in range(gd.ndims):
            for S in range(2):
                if arr[D][S].kind == 1 and arr[D][S].nfields:
                    dt = self._F(arr[D][S].fields[0]).dtype
        for D in range(gd.ndims):
            for S in range(2):
                if dt is None and arr[D][S].kind == 2 and arr[D][S].nfields:
                    dt = self._F(arr[D][S].fields[0]).dtype
        g = self._grid(gd, dt or np.float64)
        self._apply_batches(ctx, g, arr)
        self.launches += gd.ndims
        return 0

    def chmy_launch(self, ctx, dref):
        d = dref._obj
        rc = self._validate(d)
        if rc:
            return rc
        o = self.o
        nd, nt = d.grid.ndims, NT.get(d.grid.ndims, 0)
        F = [self._F(d.fields[q]) if d.fields[q] else None for q in range(d.nfields)]
        if any(f is not None and not f.has_storage for f in F):
            return self._fail(-1, "descriptor-only field")
        self._fusion_step(ctx, d, F)
        g = self._grid(d.grid, F[0].dtype)
        s = [d.scalars[q] for q in range(d.nscalars)]
        vn, tn = "xyz"[:nd], (("xx", "yy", "xy") if nd == 2 else ("xx", "yy", "zz", "xy", "xz", "yz"))
        V = lambda i: {c: F[i + k].f for k, c in enumerate(vn)}
        Tn = lambda i: {c: F[i + k].f for k, c in enumerate(tn)}
        op, args = None, None
        if d.op == 1:      # q.x q.y C ; chi
            op, args = o.compute_q, (V(0), F[2].f, s[0])
        elif d.op == 2:    # C q.x q.y ; dt
            op, args = o.update_C, (F[0].f, V(1), s[0])
        elif d.op == 3:    # T tau[nt] T_old tau_old[nt]
            op, args = o.update_old, (F[0].f, Tn(1), F[1 + nt].f, Tn(2 + nt))
        elif d.op == 4:    # tau[nt] Pr divV V[nd] tau_old[nt] ; eta eta_ve G dt dtau_Pr dtau_r
            op, args = o.update_stress, (Tn(0), F[nt].f, F[nt + 1].f, V(nt + 2), Tn(nt + 2 + nd), *s)
        elif d.op == 5:    # V[nd] r_V[nd] Pr tau[nt] rho_g|NULL ; eta_ve nudtau
            rho = F[2 * nd + 1 + nt]
            if rho is None:
                q = d.rho_g
                rho_arg = o.Inclusion(tuple(q.loc[a] for a in range(nd)), tuple(q.c0[a] for a in range(nd)), q.r, q.inn, q.out)
            else:
                rho_arg = rho.f
            op, args = o.update_velocity, (V(0), V(nd), F[2 * nd].f, Tn(2 * nd + 1), rho_arg, *s)
        elif d.op == 6:    # qT[nd] T V[nd] ; lambda
            op, args = o.update_thermal_flux, (V(0), F[nd].f, V(nd + 1), s[0])
        elif d.op == 7:    # T T_old qT[nd] ; dt
            op, args = o.update_thermal, (F[0].f, F[1].f, V(2), s[0])
        elif d.op == 8:    # operators: destinations first, then sources, then the coefficient field
            nout = nd if d.oper in (13, 14) else 1
            vec = d.oper in (9, 12)
            dst = [F[k].f for k in range(nout)]
            src = [F[nout + k].f for k in range(nd if vec else 1)]
            kf = F[nout + 1].f if d.oper in (6, 11, 14) else None
            o.apply_operator(g, d.oper, dst, src, k=kf, dim=d.oper_dim)
            self.launches += 1
            return 0
        else:
            return self._fail(-1, f"unknown op id {d.op}")
        ow = tuple(d.outer_width[a] for a in range(nd)) if d.has_outer_width else None
        if ow is not None and any(w < 3 or 2 * w > d.grid.n[a] + 2 for a, w in enumerate(ow)):
            ow = None                                            # the library's own rule: such widths do not split
        # the ops are pointwise writers: regions + per-dim batches in the reference's order == full range, then the batches
        la = o.Launcher(g, ow)
        o.launch(la, g, op, args, bc=None)
        if d.has_bc:
            self._apply_batches(ctx, g, d.bc)
        self.launches += 1 + (nd if d.has_bc else 0)
        return 0
