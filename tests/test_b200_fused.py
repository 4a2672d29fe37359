"""
GPU parity of the lazily fused update_stress! + update_velocity! sweep (ops_fused.cu, chmy_set_fusion) through the C
ABI: (1) one iteration on random fields (padding included) against the oracle's two kernels + boundary batch, for
every tile geometry (rows per CTA, cluster size, z-chunk), both rho_g sources and both division paths, bit-exact on
the full padded arrays; (2) the deferred launch is executed before anything can observe it; (3) whole solver runs
(ping-pong over many iterations, thermal sub-steps and residual checks in between, split launches on two streams)
are bit-identical to the two-kernel path.  The same phase functions are proven on the CPU by
tests/test_fused_emulation.py; this file proves the compiled kernel, the cluster / DSMEM exchange and the host glue.
"""
import numpy as np
import pytest

from helpers import assert_same, fill_pair
from test_b200_parity import _set_tuning, _stokes_pair, bloc, mk_grids  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    return chmy_b200


@pytest.fixture()
def arch(ch):
    a = ch.Arch(ch.B200Backend())
    ch.set_fusion(a, True)
    yield a
    a.close()                      # the tile geometry is per context: nothing to restore


SC = dict(eta=10.0, eta_ve=0.737, G=1.3, dt=0.0171, dPr=0.0213, dr=0.613, nud=0.00931)


def _bc_V(ch_or_o, V, names=("x", "y", "z")):
    D, N = ch_or_o.Dirichlet, ch_or_o.Neumann
    comps = [V[c] for c in names] if isinstance(V, dict) else list(V)
    return tuple((comps[i], {a: (D() if a == names[i] else N()) for a in names}) for i in range(3))


GEOMS = [(8, 4, 64), (4, 1, 3), (8, 1, 5), (6, 4, 64), (4, 8, 7), (6, 3, 4), (8, 8, 16), (4, 2, 1), (6, 5, 9), (6, 1, 2)]


@pytest.mark.parametrize("n", [(70, 37, 9), (17, 9, 5), (125, 64, 20), (1, 1, 1), (60, 6, 3)])
@pytest.mark.parametrize("geom", GEOMS)
def test_fused_iteration_bit_exact_vs_oracle(ch, arch, oracle, n, geom):
    o = oracle
    tyb, cl, cz = geom
    ch.set_fused_tuning(arch, tyb, cl, cz, (tyb + cz + n[0]) % 2)
    fun = (sum(n) + tyb) % 2 == 0
    _set_tuning(disable_fast=0, true_div=(cl + cz) % 2)
    rng = np.random.default_rng(7 + cz)
    og, bg, O, B, pairs = _stokes_pair(ch, o, arch, n, rng)
    Lo, Lb = o.Launcher(og), ch.Launcher(arch, bg)
    if fun:
        rho_o = o.Inclusion((0, 0, 1), (0.02, -0.01, 0.03), 0.33, 1.0, 0.0)
        rho_b = ch.FunctionField(ch.init_incl, bg, bloc(ch, (0, 0, 1)),
                                 parameters={"x0": 0.02, "y0": -0.01, "z0": 0.03, "r": 0.33, "in": 1.0, "out": 0.0})
    else:
        rho_o, rho_b = O["rho"], B["rho"]
    n0 = ch.fused_count(arch)
    for it in range(3):                       # three sweeps: both ping-pong parities and a re-used shadow
        o.launch(Lo, og, o.update_stress, (O["tau"], O["Pr"], O["dV"], O["V"], O["tau_old"], SC["eta"], SC["eta_ve"],
                                           SC["G"], SC["dt"], SC["dPr"], SC["dr"]))
        o.launch(Lo, og, o.update_velocity, (O["V"], O["rV"], O["Pr"], O["tau"], rho_o, SC["eta_ve"], SC["nud"]),
                 bc=o.batch(og, *_bc_V(o, O["V"])))
        Lb(arch, bg, (ch.update_stress_, (B["tau"], B["Pr"], B["dV"], B["V"], B["tau_old"], SC["eta"], SC["eta_ve"],
                                          SC["G"], SC["dt"], SC["dPr"], SC["dr"], bg)))
        Lb(arch, bg, (ch.update_velocity_, (B["V"], B["rV"], B["Pr"], B["tau"], rho_b, SC["eta_ve"], SC["nud"], bg)),
           bc=ch.batch(bg, *_bc_V(ch, B["V"])))
        assert ch.fused_count(arch) == n0 + it + 1, "the pair of launches did not take the fused path"
        # division mode of the sweep: div.rn.f64 on request, else two operations when proven exact for all four divisors
        two = all(ch.division_two_op_exact(c) for c in (SC["G"] * SC["dt"], SC["eta"], SC["eta_ve"], 3.0))
        assert ch.last_division_mode(arch) == (1 if (cl + cz) % 2 else (2 if two else 0))
        for name, a, b in pairs:
            assert_same(a, b, f"sweep {it}: {name}")
    _set_tuning(0, 0)


DRIVER_GEOMS = [(6, 4, 64, 1), (4, 4, 64, 0), (6, 5, 7, 1), (8, 2, 5, 1), (4, 6, 64, 1), (6, 1, 3, 0)]


@pytest.mark.parametrize("geom", DRIVER_GEOMS)
@pytest.mark.parametrize("n", [(70, 37, 9), (125, 64, 20), (17, 9, 5)])
@pytest.mark.parametrize("overlap", [True, False])
def test_fused_solver_bit_exact_vs_two_kernels_any_geometry(ch, n, geom, overlap):
    """the whole driver (batches folded into one launch, boundary tiles first + retire counter when `overlap`) against the two
    tuned kernels with one launch per dimension on one stream: same bits in every field"""
    from chmy_b200 import drivers as BD
    res = []
    for fused in (True, False):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, fused)
        ch.set_fused_tuning(a, *geom)
        ch.set_launch_split(a, "always" if (fused and overlap) else False, bc_fold=fused)
        s = BD.Stokes(a, n, rho_g_function=(n[0] % 2 == 0), blocking=False)
        rng = np.random.default_rng(1)
        for f in s.fields().values():
            f.from_host(1e-3 * (rng.random(tuple(d + 4 for d in f.dims)) - 0.5), [-1] * 3, [d + 2 for d in f.dims])
        s.run(1, 7, 7, eps=0.0)
        if fused:
            assert ch.fused_count(a) == 7 and ch.overlapped_count(a) == (7 if overlap else 0)
        res.append({k: f.parent() for k, f in s.fields().items()})
        a.close()
    for k in res[0]:
        assert ((res[0][k] == res[1][k]) | (np.isnan(res[0][k]) & np.isnan(res[1][k]))).all(), k


def test_deferred_stress_is_flushed_before_it_can_be_observed(ch, arch, oracle):
    o = oracle
    n = (33, 18, 7)
    rng = np.random.default_rng(3)
    og, bg, O, B, pairs = _stokes_pair(ch, o, arch, n, rng)
    Lo, Lb = o.Launcher(og), ch.Launcher(arch, bg)
    n0 = ch.fused_count(arch)
    o.launch(Lo, og, o.update_stress, (O["tau"], O["Pr"], O["dV"], O["V"], O["tau_old"], SC["eta"], SC["eta_ve"],
                                       SC["G"], SC["dt"], SC["dPr"], SC["dr"]))
    Lb(arch, bg, (ch.update_stress_, (B["tau"], B["Pr"], B["dV"], B["V"], B["tau_old"], SC["eta"], SC["eta_ve"],
                                      SC["G"], SC["dt"], SC["dPr"], SC["dr"], bg)))
    for name, a, b in pairs:                  # copy_to_host must see the stress update
        assert_same(a, b, "flush: " + name)
    assert ch.fused_count(arch) == n0
    # a velocity launch on OTHER fields must not fuse with a pending stress launch
    og2, bg2, O2, B2, pairs2 = _stokes_pair(ch, o, arch, n, rng)
    Lb(arch, bg, (ch.update_stress_, (B["tau"], B["Pr"], B["dV"], B["V"], B["tau_old"], SC["eta"], SC["eta_ve"],
                                      SC["G"], SC["dt"], SC["dPr"], SC["dr"], bg)))
    Lb(arch, bg2, (ch.update_velocity_, (B2["V"], B2["rV"], B2["Pr"], B2["tau"], B2["rho"], SC["eta_ve"], SC["nud"], bg2)))
    o.launch(Lo, og, o.update_stress, (O["tau"], O["Pr"], O["dV"], O["V"], O["tau_old"], SC["eta"], SC["eta_ve"],
                                       SC["G"], SC["dt"], SC["dPr"], SC["dr"]))
    o.launch(o.Launcher(og2), og2, o.update_velocity, (O2["V"], O2["rV"], O2["Pr"], O2["tau"], O2["rho"], SC["eta_ve"], SC["nud"]))
    assert ch.fused_count(arch) == n0
    for name, a, b in pairs + pairs2:
        assert_same(a, b, "mismatched pair: " + name)


@pytest.mark.parametrize("n,ow,exact", [((70, 37, 20), None, False), ((130, 40, 36), (8, 4, 3), True),
                                        ((64, 64, 64), (16, 8, 4), True), ((63, 9, 5), None, False)])
def test_fused_solver_equals_two_kernel_solver(ch, oracle, n, ow, exact):
    """2 outer steps x 30 PT iterations with thermal sub-steps and residual checks: fused == unfused, every bit of every
    padded array and the residual history (so the ping-pong swap, the frame carry-over and the flushes are right)."""
    from chmy_b200 import drivers as BD
    res, hist = [], []
    for fused in (True, False):
        a = ch.Arch(ch.B200Backend())
        ch.set_fusion(a, fused)
        s = BD.Stokes(a, n, rho_g_function=(n[0] % 2 == 0), outer_width=ow, exact_split=exact, blocking=not exact)
        rng = np.random.default_rng(42)
        for f in s.fields().values():
            f.from_host(1e-3 * (rng.random(tuple(d + 4 for d in f.dims)) - 0.5), [-1] * 3, [d + 2 for d in f.dims])
        c0 = ch.fused_count(a)
        hist.append(s.run(2, 30, 10, eps=0.0))
        assert (ch.fused_count(a) - c0 == 60 + 30) if fused else (ch.fused_count(a) == c0)   # 60 mechanics + 30 thermal sweeps
        res.append({k: f.parent() for k, f in s.fields().items()})
        a.close()
    assert hist[0] == hist[1]
    for k in res[0]:
        same = (res[0][k] == res[1][k]) | (np.isnan(res[0][k]) & np.isnan(res[1][k]))
        assert same.all(), k


def test_fused_solver_matches_oracle(ch, arch, oracle):
    import drivers as OD
    from chmy_b200 import drivers as BD
    n = (30, 22, 14)
    osol = OD.Stokes(n, rho_g_function=True)
    bsol = BD.Stokes(arch, n, rho_g_function=True)
    ho, hb = osol.run(2, 110, 22), bsol.run(2, 110, 22)
    assert ch.fused_count(arch) >= 220
    assert len(ho) == len(hb) == 10
    for a, b in zip(ho, hb):
        assert a[:2] == b[:2]
        for x, y in zip(a[2:], b[2:]):
            assert abs(x - y) <= 1e-12 * abs(x), (a, b)
    bf = bsol.fields()
    for k, f in osol.fields().items():
        assert_same(f, bf[k], k, tol=1e-12)
