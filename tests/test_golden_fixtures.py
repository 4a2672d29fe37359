"""
Committed fixtures under tests/golden/ (made by tests/golden/make_golden.py from the oracle; regression vectors, not
reference outputs -- the reference is Julia and cannot run in this image).

  not gpu : the oracle, rebuilt on this host, must reproduce the stored bits exactly (pins the checker itself);
  gpu     : the CUDA path through the C ABI must match the stored arrays (<= 1e-12 relative on full padded arrays and
            residual histories; bit-exact for halo pack buffers).
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_oracle_reproduces_golden_bits(oracle, name):
    kind, n, kw = MG.CASES[name]
    got, want = MG.run_case(kind, n, kw), _load(name)
    assert sorted(got) == sorted(want)
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), f"{name}:{k}"


def test_oracle_reproduces_halo_pack_bits(oracle):
    got, want = MG.halo_case(), _load("halo_pack")
    assert sorted(got) == sorted(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


# ------------------------------------------------------------------------------------------------ CUDA path vs stored bits
def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))


@pytest.fixture(scope="module")
def arch():
    import chmy_b200 as ch
    a = ch.Arch(ch.B200Backend())
    yield a
    a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", [k for k, v in sorted(MG.CASES.items()) if v[0] in ("stokes", "diffusion")])
def test_cuda_matches_golden(arch, name):
    from chmy_b200 import drivers as BD
    kind, n, kw = MG.CASES[name]
    want = _load(name)
    if kind == "stokes":
        s = BD.Stokes(arch, n, rho_g_function=kw["rho_g_function"])
        h = np.array(s.run(kw["nt"], kw["niter"], kw["ncheck"]), dtype=np.float64)
        assert h.shape == want["history"].shape
        assert np.array_equal(h[:, :2], want["history"][:, :2])
        assert np.all(np.abs(h[:, 2:] - want["history"][:, 2:]) <= 1e-12 * np.abs(want["history"][:, 2:]))
        assert s.dt == want["dt_eta_ve"][0] and s.eta_ve == want["dt_eta_ve"][1]
    else:
        s = BD.Diffusion2D(arch, n, C0=np.random.default_rng(kw["seed"]).random(n))
        s.run(kw["nt"])
    for k, f in s.fields().items():
        a = want["f:" + k]
        assert _rel(a, f.parent()) <= 1e-12, f"{name}:{k}"


@pytest.mark.gpu
def test_cuda_halo_pack_matches_golden(arch):
    import ctypes as C
    import chmy_b200 as ch
    from chmy_b200 import _lib as L
    want = _load("halo_pack")
    for n, loc in (((9, 6), (1, 0)), ((7, 5, 4), (0, 1, 1)), ((7, 5, 4), (1, 0, 0))):
        g = ch.UniformGrid(arch, origin=(-1.0,) * len(n), extent=(2.0,) * len(n), dims=n)
        f = ch.Field(arch, g, tuple(ch.Vertex() if x else ch.Center() for x in loc))
        sd = tuple(d + 4 for d in f.dims)
        f.from_host(np.arange(int(np.prod(sd)), dtype=np.float64).reshape(sd, order="F"), [-1] * len(n), [d + 2 for d in f.dims])
        tag = "x".join(map(str, n)) + "_" + "".join(map(str, loc))
        for D in range(len(n)):
            for S in range(2):
                ref = want[f"pack:{tag}:{D}{S}"]
                buf = np.empty(ref.size, dtype=np.float64)
                L.check(L.lib().chmy_halo_pack(arch.ctx, f.handle, D, S, buf.ctypes.data_as(C.c_void_p)))
                assert np.array_equal(buf, ref.reshape(-1, order="F")), (tag, D, S)
