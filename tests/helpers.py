"""Shared helpers of the parity tests: seeded inputs pushed identically into the oracle and the CUDA path."""
import ctypes
import glob
import os
import subprocess

import numpy as np

_EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "chmy.jl_b200", "csrc")


def build_emul(name):
    """Compile tests/emul/<name>.cpp (a host emulation of a kernel: it #includes the kernel's own .cuh) and load it.
    CHMY_EMUL_SANITIZE=1 builds with AddressSanitizer + UBSan instead (tests/test_emulation_sanitizers.py runs the
    emulation suites that way, in a child process that preloads the sanitizer runtimes): every load / store of the
    kernels' phase functions is then checked against the bounds of the PITCHED arrays the tests hand them."""
    san = os.environ.get("CHMY_EMUL_SANITIZE", "0") == "1"
    src = os.path.join(_EMUL, name + ".cpp")
    lib = os.path.join(_EMUL, f"lib{name}{'_san' if san else ''}.so")
    deps = [src] + glob.glob(os.path.join(_CSRC, "*.cuh"))
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if san else ["-O2"]
        import platform
        # x86-64: AVX2 + FMA so that fma() is one correctly rounded instruction; elsewhere the toolchain's default (aarch64 has fma)
        arch = ["-march=x86-64-v3"] if platform.machine() in ("x86_64", "AMD64") else []
        subprocess.check_call(["g++"] + flags + ["-ffp-contract=off"] + arch + ["-shared", "-fPIC", "-Wall",
                                                 "-Wno-unknown-pragmas", "-Wno-unused-function", "-o", lib, src])
    return ctypes.CDLL(lib)


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


def install_dryrun_if_requested():
    """Child processes of the dry run (tests/test_gpu_suite_dryrun.py) that import chmy_b200 on their own -- the ranks the
    multi-GPU suite spawns -- put tests/dryrun_backend.py in the library's place.  No-op unless CHMY_DRYRUN=1."""
    if os.environ.get("CHMY_DRYRUN") != "1":
        return False
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import chmy_b200
    from chmy_b200 import _lib
    import oracle as o
    from dryrun_backend import DryRunLib
    if not isinstance(_lib.lib(), DryRunLib):
        fake = DryRunLib(_lib.lib(), o)
        _lib.lib = lambda: fake
        chmy_b200.load_library = _lib.lib
    return True
