"""
Parity at BASELINE.json's FULL sizes (GPU only), through size-independent properties -- the oracle cannot hold 767^3.

The generic kernels (one thread per cell, any box, any layout; chmy.jl_b200/csrc/ops.cu) are proven bit-identical to the
oracle at every size the oracle finishes in seconds (tests/test_b200_parity.py), and their per-cell arithmetic does not
depend on the grid size.  At the headline sizes the product paths must therefore reproduce them bit for bit:

    fused sweep (bench default)  ==  two tuned kernels  ==  generic kernels        3D Stokes 767^3   (config 4/5)
    tuned kernels                ==  generic kernels                               2D Stokes 8191^2  (config 3)
    tuned kernels                ==  generic kernels                               2D diffusion 16383^2 (config 2)

compared through exact checksums (wrapping uint64 sum and xor of the bit patterns) plus max|f| over the WHOLE field from
the device reduction, after a few PT iterations that include the boundary batches.  2D fields are checksummed whole;
the 3.7 GB 3D fields through full xy-planes and full xz-planes at indices that straddle every kind of tile edge of the
kernels (first/last cells, 60-cell row segments, cluster rows, 64-plane z-chunks, the middle, both halos) -- every x
index, every y index and every z index of the launch range is covered by some probe.  A literal split launch
(outer_width honoured, two streams) must give the same checksums as the single full-range launch.

Green on the driver's B200 at the end of round 1; enforced (no xfail) since round 2.
"""
import gc
import math
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


# tests/test_gpu_suite_dryrun.py executes this file on the CPU against the dry-run backend: same code, sizes the oracle holds
DRY = os.environ.get("CHMY_DRYRUN") == "1"
N3 = (40, 36, 28) if DRY else (767, 767, 767)
N2S = (96, 80) if DRY else (8191, 8191)
N2D = (128, 96) if DRY else (16383, 16383)
OW3 = (8, 4, 3) if DRY else (128, 8, 4)


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    return chmy_b200


def _set_tuning(disable_fast=-1, true_div=-1):
    from chmy_b200 import _lib as L
    L.check(L.lib().chmy_set_tuning(disable_fast, true_div))


def _probe_boxes(dims):
    """boxes (lo, hi) in logical indices, halo included: everything for 1D/2D fields, probe planes for 3D fields"""
    nd = len(dims)
    lo, hi = [0] * nd, [d + 1 for d in dims]
    if nd < 3 or int(np.prod(dims)) < (1 << 24):
        return [(lo, hi)]
    boxes = []
    for axis, picks in ((2, (0, 1, 2, 63, 64, 65, 128)), (1, (0, 1, 2, 13, 14, 15, 28))):
        d = dims[axis]
        for i in sorted(set(p for p in picks + (d // 2, d - 1, d, d + 1) if 0 <= p <= d + 1)):
            l, h = list(lo), list(hi)
            l[axis] = h[axis] = i
            boxes.append((l, h))
    return boxes


def checksums(ch, fields):
    """{name: (uint64 wrapping sum, uint64 xor, max|f| over the whole field)}, one probe on the host at a time"""
    out = {}
    for k, f in fields.items():
        s, x = 0, 0
        for lo, hi in _probe_boxes(f.dims):
            bits = f.to_host(lo, hi).reshape(-1, order="F").view(np.uint64)
            s = (s + int(bits.sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
            x ^= int(np.bitwise_xor.reduce(bits))
            del bits
        out[k] = (s, x, ch.maxabs(f, with_halo=True))
    return out


def run_stokes(ch, n, iters, *, fused, generic, outer_width=None, exact_split=False):
    from chmy_b200 import drivers as BD
    arch = ch.Arch(ch.B200Backend())
    try:
        _set_tuning(disable_fast=1 if generic else 0)
        ch.set_fusion(arch, bool(fused))
        sol = BD.Stokes(arch, n, re_m=2.5 * math.pi, rho_g_function=True, outer_width=outer_width, adv_coef=0.01,
                        blocking=False, exact_split=exact_split)
        sol.begin_time_step()
        for _ in range(iters):
            sol.mechanics()
        sol.thermal()                                   # one thermal sub-step: flux + update + T batch
        res = sol.residuals()
        ch.synchronize(arch)
        nf = ch.fused_count(arch)
        cs = checksums(ch, sol.fields())
        for f in sol.fields().values():
            f.free()
        del sol
        return cs, res, nf
    finally:
        _set_tuning(disable_fast=0)
        arch.close()
        gc.collect()


def same(a, b, what):
    assert a.keys() == b.keys()
    bad = [k for k in a if a[k][:2] != b[k][:2] or not (a[k][2] == b[k][2] or (math.isnan(a[k][2]) and math.isnan(b[k][2])))]
    assert not bad, f"{what}: fields differ: {bad}"
    assert any(v[2] > 0 for v in a.values())


def test_stokes3d_767_fused_equals_tuned_equals_generic(ch):
    n, iters = N3, 4
    cs_f, res_f, nf = run_stokes(ch, n, iters, fused=True, generic=False)
    assert nf == iters + 1                               # the bench path really ran: one sweep per PT iteration + the thermal sweep
    cs_t, res_t, _ = run_stokes(ch, n, iters, fused=False, generic=False)
    same(cs_f, cs_t, "fused sweep vs two tuned kernels at 767^3")
    assert res_f == res_t
    cs_g, res_g, _ = run_stokes(ch, n, iters, fused=False, generic=True)
    same(cs_t, cs_g, "tuned kernels vs generic kernels at 767^3")
    assert res_t == res_g and all(math.isfinite(r) for r in res_g)


def test_stokes3d_767_split_launch_equals_full_range(ch):
    """Launcher(outer_width=(128, 8, 4)) honoured literally (inner region + 6 slabs on two streams, KernelLaunch.jl:160-181)
    vs the single full-range launch, at the headline size, for the fused and the two-kernel path."""
    n, iters = N3, 3
    for fused in (True, False):
        a, ra, _ = run_stokes(ch, n, iters, fused=fused, generic=False)
        b, rb, _ = run_stokes(ch, n, iters, fused=fused, generic=False, outer_width=OW3, exact_split=True)
        same(a, b, f"split vs unsplit at 767^3 (fused={fused})")
        assert ra == rb


def test_stokes2d_8191_tuned_equals_generic(ch):
    n, iters = N2S, 6
    a, ra, _ = run_stokes(ch, n, iters, fused=False, generic=False)
    b, rb, _ = run_stokes(ch, n, iters, fused=False, generic=True)
    same(a, b, "2D Stokes tuned vs generic at 8191^2")
    assert ra == rb


def test_diffusion2d_16383_tuned_equals_generic(ch):
    from chmy_b200 import drivers as BD
    n, iters = N2D, 5
    out = []
    for generic in (False, True):
        arch = ch.Arch(ch.B200Backend())
        try:
            _set_tuning(disable_fast=1 if generic else 0)
            sol = BD.Diffusion2D(arch, n, outer_width=(16, 8) if DRY else (128, 8), C0=None, blocking=False)
            rng = np.random.default_rng(0)
            for j0 in range(1, n[1] + 1, 2048):                           # rand() initial condition, uploaded in strips
                j1 = min(n[1], j0 + 2047)
                sol.C.from_host(rng.random((n[0], j1 - j0 + 1)), [1, j0], [n[0], j1])
            ch.bc_(arch, sol.grid, (sol.C, ch.Neumann()), exchange=sol.C)
            sol.run(iters)
            out.append(checksums(ch, sol.fields()))
            for f in sol.fields().values():
                f.free()
            del sol
        finally:
            _set_tuning(disable_fast=0)
            arch.close()
            gc.collect()
    same(out[0], out[1], "diffusion tuned vs generic at 16383^2")
