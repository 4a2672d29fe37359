"""
CPU-only: the CPU oracle (both element-type builds) and the host emulations of the CUDA kernels (fused stress+velocity sweep, 2D sweeps, flux->update pairs, 3D thermal
sweep, grid operators -- the kernels' own .cuh sources compiled by g++) run once more under AddressSanitizer + UBSan.
The emulation suites hand the kernels numpy arrays in the library's PITCHED layout (same pitch, lead-in and slack as
chmy_field_create allocates on the device), so a load or store outside a field's allocation -- which a GPU would turn
into a silent read of a neighbouring buffer -- is reported here, for every tile geometry, sub-box and grid size the
suites cover.  Runs in a child process that preloads the sanitizer runtimes.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = ["tests/test_fused_emulation.py", "tests/test_fused_emulation2d.py", "tests/test_fused_emulation_t3.py",
          "tests/test_operators_emulation.py",
          # ... and the oracle itself (the parity anchor) under the same sanitizers, through the suites that pin it
          "tests/test_oracle_golden.py", "tests/test_oracle_transliteration.py", "tests/test_host_transliteration.py",
          "tests/test_golden_fixtures.py"]


def _runtime(name):
    out = subprocess.run(["g++", f"-print-file-name={name}"], capture_output=True, text=True).stdout.strip()
    return out if os.path.isabs(out) and os.path.exists(out) else None


def test_kernel_emulations_are_clean_under_asan_and_ubsan():
    asan, ubsan = _runtime("libasan.so"), _runtime("libubsan.so")
    if not asan or not ubsan:
        pytest.skip("sanitizer runtimes not installed")
    env = dict(os.environ, CHMY_EMUL_SANITIZE="1", CHMY_ORACLE_SANITIZE="1", OMP_NUM_THREADS="2", LD_PRELOAD=f"{asan}:{ubsan}",
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + SUITES, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "AddressSanitizer" not in tail and "runtime error" not in tail, tail
    assert " passed" in r.stdout
