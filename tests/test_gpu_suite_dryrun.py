"""
CPU-only: the single-GPU `-m gpu` test files, EXECUTED on the CPU against tests/dryrun_backend.py.

Most of what can break a GPU test file is not CUDA: the test's own Python, the host mirror's flattening of fields, batches
and descriptors, element-type plumbing, the drivers, the ABI's argument rules.  The dry-run backend restates the C ABI over
the CPU oracle (descriptors decoded by the field order include/chmy_b200.h documents, arguments validated by the REAL
library on descriptor-only twins of the fields), so a child pytest process can run those files here.  Such a run compares
the oracle with itself: it says NOTHING about the CUDA kernels and is not parity evidence -- it says that the GPU suites
are runnable programs whose host side does what the header says, before a GPU minute is spent on them.

Deselected in the dry run: tests that measure a device property (alignment of real allocations, the division self-test,
launch / fused-sweep counters, full BASELINE sizes) or need real device pointers.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = ["tests/test_b200_parity.py", "tests/test_golden_fixtures.py", "tests/test_b200_fused.py", "tests/test_b200_fused2d.py",
         "tests/test_zy_b200_fullsize.py", "tests/test_zz_b200_round2.py"]
# not in the dry run: the multi-GPU suite (needs ranks) and the device self-test of the exact-division sequence.  The
# fused-sweep suites run with their gated (CHMY_EXPERIMENTAL) cases; the full-size suite runs at sizes the oracle can hold.
DESELECT = "not test_exact_division_by_uniform_scalar"


def test_single_gpu_suites_run_on_the_dry_run_backend():
    env = dict(os.environ, CHMY_DRYRUN="1", CHMY_EXPERIMENTAL="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "--runxfail", "-p", "no:cacheprovider", "-k", DESELECT] + FILES,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1700)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    last = r.stdout.strip().splitlines()[-1]
    assert " passed" in last and "failed" not in last and "error" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 100, last          # the files really ran (not everything deselected)
