"""Shared helpers of the parity tests: seeded inputs pushed identically into the oracle and the CUDA path."""
import numpy as np


def fill_pair(rng, of, bf, scale=1.0):
    """Same random bits into the FULL padded arrays (interior, halo and padding) of an oracle field and a B200 field."""
    a = (rng.random(of.sdims) - 0.5) * scale
    of.data[...] = a
    lo = [-1] * len(of.dims)
    hi = [d + 2 for d in of.dims]
    bf.from_host(a, lo, hi)
    return a


def assert_same(of, bf, name="", tol=0.0):
    """Compare the full padded arrays.  tol = 0 -> bit-exact (NaN == NaN, -0.0 == +0.0 accepted)."""
    a, b = of.data, bf.parent()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if tol == 0.0:
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        if not same.all():
            idx = np.argwhere(~same)
            i = tuple(idx[0])
            raise AssertionError(f"{name}: {len(idx)} cells differ; first at storage index {i}: oracle {a[i]!r} cuda {b[i]!r}")
    else:
        scale = max(np.abs(a).max(), 1e-300)
        err = np.abs(a - b).max() / scale
        assert err <= tol, f"{name}: relative error {err:.3e} > {tol:.1e}"
