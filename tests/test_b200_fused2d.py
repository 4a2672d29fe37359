"""
GPU parity of the 2D / thermal fused sweeps (ops_fused2d.cu, chmy_set_fusion(ctx, 3)) through the C ABI:
update_stress! + update_velocity! (2D), compute_q! + update_C!, update_thermal_flux! + update_thermal! (2D).

Whole solver runs (ping-pong over many iterations, boundary batches and residual checks in between, literal split
launches on two streams) must be bit-identical to the two-kernel path and agree with the oracle's drivers, for every
chunk / load-group setting.  The phase functions are proven on the CPU by tests/test_fused_emulation2d.py; this file
proves the compiled kernels and the host glue.

First run on a B200 in round 2 (profiles/r2_c1_gpu_tests.log: all green); part of the enforced suite since.
"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    return chmy_b200


def _same(a, b, name):
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    if not ok.all():
        i = tuple(np.argwhere(~ok)[0])
        raise AssertionError(f"{name}: {int((~ok).sum())} cells differ; first at storage {i}: two-kernel {a[i]!r} fused {b[i]!r}")


TUNINGS = [(64, 4), (1, 1), (3, 2), (128, 4), (7, 1)]


def _odd(n, ow, exact):
    """a literal split with odd x slab starts takes the two-kernel path (api.cu: odd_exact_split)"""
    return bool(exact and ow is not None and ((ow[0] & 1) or ((n[0] + 2 - ow[0]) & 1)))


@pytest.mark.parametrize("tuning", TUNINGS)
@pytest.mark.parametrize("n,ow,exact", [((70, 37), (16, 8), False), ((126, 64), (16, 8), True), ((17, 9), (4, 3), True),
                                        ((256, 256), (128, 8), False), ((62, 130), (6, 4), True)])
def test_fused_diffusion_equals_two_kernels_and_oracle(ch, oracle, n, ow, exact, tuning):
    from chmy_b200 import drivers as BD
    import drivers as OD
    C0 = np.random.default_rng(5).random(n)
    res = []
    for mode in (3, 0):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, mode)
        ch.set_fused2d_tuning(a, *tuning)
        s = BD.Diffusion2D(a, n, outer_width=ow, C0=C0, exact_split=exact)
        s.run(9)
        if mode:
            assert ch.fused_count(a) == (0 if _odd(n, ow, exact) else 9), ch.fused_count(a)
        res.append({k: f.parent() for k, f in s.fields().items()})
        a.close()
    for k in res[0]:
        _same(res[1][k], res[0][k], k)
    o = OD.Diffusion2D(n, outer_width=ow, C0=C0)
    o.run(9)
    for k, f in o.fields().items():
        _same(f.data, res[0][k], "oracle " + k)


@pytest.mark.parametrize("tuning", TUNINGS[:3])
@pytest.mark.parametrize("n,fun,ow,exact", [((70, 37), True, None, False), ((126, 64), False, (16, 8), True), ((125, 64), False, (16, 8), True),
                                            ((17, 9), False, None, False), ((130, 61), True, (6, 4), True)])
def test_fused_stokes2d_with_thermal_equals_two_kernels(ch, oracle, n, fun, ow, exact, tuning):
    from chmy_b200 import drivers as BD
    import drivers as OD
    res, hist = [], []
    for mode in (3, 0):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, mode)
        ch.set_fused2d_tuning(a, *tuning)
        s = BD.Stokes(a, n, rho_g_function=fun, outer_width=ow, exact_split=exact)
        hist.append(s.run(2, 12, 6, eps=0.0))
        if mode:
            # 24 mechanics sweeps + 12 thermal sweeps (it = 2)
            assert ch.fused_count(a) == (0 if _odd(n, ow, exact) else 24 + 12), ch.fused_count(a)
        res.append({k: f.parent() for k, f in s.fields().items()})
        a.close()
    assert hist[0] == hist[1]
    for k in res[0]:
        _same(res[1][k], res[0][k], k)
    o = OD.Stokes(n, rho_g_function=fun, outer_width=ow)
    ho = o.run(2, 12, 6, eps=0.0)
    assert len(ho) == len(hist[0])
    for x, y in zip(ho, hist[0]):
        assert x[:2] == y[:2] and np.allclose(x[2:], y[2:], rtol=1e-12, atol=0.0)
    for k, f in o.fields().items():
        err = float(np.abs(f.data - res[0][k]).max() / max(np.abs(f.data).max(), 1e-300))
        assert err <= 1e-12, (k, err)


def test_deferred_2d_launch_is_flushed_before_it_can_be_observed(ch, oracle):
    """compute_q! is deferred under mode 3; reading q must see it executed (two-kernel fallback), and the following
    update_C! then runs on its own."""
    from chmy_b200 import drivers as BD
    n = (33, 18)
    C0 = np.random.default_rng(2).random(n)
    outs = []
    for mode in (3, 0):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, mode)
        s = BD.Diffusion2D(a, n, C0=C0)
        s.launch(a, s.grid, (ch.compute_q_, (s.q, s.C, s.chi, s.grid)))
        qx = s.q.x.parent()                       # observes q: the deferred launch must have run
        s.launch(a, s.grid, (ch.update_C_, (s.C, s.q, s.dt, s.grid)), bc=ch.batch(s.grid, (s.C, ch.Neumann()), exchange=s.C))
        ch.synchronize(a)
        assert ch.fused_count(a) == 0
        outs.append((qx, s.C.parent()))
        a.close()
    _same(outs[1][0], outs[0][0], "q.x")
    _same(outs[1][1], outs[0][1], "C")


# ------------------------------------------------------------------------------------------------ 3D thermal pair (kind 4)
@pytest.mark.parametrize("n,fun,ow,exact", [((70, 37, 9), True, None, False), ((126, 40, 20), False, (16, 8, 4), True),
                                            ((17, 9, 5), False, None, False), ((130, 24, 11), True, (6, 4, 3), True),
                                            ((125, 30, 12), False, (8, 4, 3), True)])
def test_fused_thermal3_equals_two_kernels_and_oracle(ch, oracle, n, fun, ow, exact):
    """3D Stokes + thermal sub-step with chmy_set_fusion(ctx, 3): the mechanics pair runs as k_fused_sv (measured path) and
    update_thermal_flux! + update_thermal! as k_fused_t3 (fused_thermal3.cuh; proven by tests/test_fused_emulation_t3.py).
    Fields and residual histories must be bit-identical to the two-kernel path and agree with the oracle's driver."""
    from chmy_b200 import drivers as BD
    import drivers as OD
    res, hist = [], []
    for mode in (3, 0):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, mode)
        s = BD.Stokes(a, n, rho_g_function=fun, outer_width=ow, exact_split=exact)
        hist.append(s.run(2, 10, 5, eps=0.0))
        if mode:
            # 20 mechanics sweeps + 10 thermal sweeps (it = 2); odd literal splits take the two-kernel path
            assert ch.fused_count(a) == (0 if _odd(n, ow, exact) else 20 + 10), ch.fused_count(a)
        res.append({k: f.parent() for k, f in s.fields().items()})
        a.close()
    assert hist[0] == hist[1]
    for k in res[0]:
        _same(res[1][k], res[0][k], k)
    o = OD.Stokes(n, rho_g_function=fun, outer_width=ow)
    ho = o.run(2, 10, 5, eps=0.0)
    assert len(ho) == len(hist[0])
    for x, y in zip(ho, hist[0]):
        assert x[:2] == y[:2] and np.allclose(x[2:], y[2:], rtol=1e-12, atol=0.0)
    for k, f in o.fields().items():
        err = float(np.abs(f.data - res[0][k]).max() / max(np.abs(f.data).max(), 1e-300))
        assert err <= 1e-12, (k, err)
