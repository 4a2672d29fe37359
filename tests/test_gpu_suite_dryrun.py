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


_CHILD = r"""
import os, sys, json, io
from contextlib import redirect_stdout
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import chmy_b200
from chmy_b200 import _lib
import oracle as o
from dryrun_backend import DryRunLib
fake = DryRunLib(_lib.lib(), o)
_lib.lib = lambda: fake
chmy_b200.load_library = _lib.lib
import __graft_entry__ as g
g.smoke()                                            # the driver's smoke(): fused 3D Stokes vs the oracle, sweep counter
import importlib.util
spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
for wl, n in (("stokes3d", "20 16 12"), ("stokes3d_thermal", "20 16 12"), ("stokes2d", "40 30"), ("diffusion2d", "40 30")):
    for fused in ("1", "3"):
        sys.argv = ["bench.py", "--workload", wl, "--n", *n.split(), "--steps", "4", "--warmup", "3", "--fused", fused, "--no-cpu-baseline"]
        buf = io.StringIO()
        with redirect_stdout(buf):
            bench.run_b200(bench.parse())
        j = json.loads(buf.getvalue().strip().splitlines()[-1])
        assert j["value"] > 0 and "host_segment_error" not in j["e2e"], j["e2e"]
        assert j["e2e"]["h2d_bytes_per_step"] > 1000 and j["roofline"]["frac"] > 0
        print("bench ok", wl, fused, j["fused_sweeps"], j["gpu_launches"])
"""


def test_smoke_and_bench_run_on_the_dry_run_backend():
    """__graft_entry__.smoke() and bench.py's B200 arm (every workload, default and experimental fusion modes) executed end
    to end -- real drivers, real host mirror, real descriptor validation -- with the dry-run backend in place of the device.
    Timings and throughputs printed by such a run are meaningless; the point is that the scripts the driver runs on the GPU
    box are runnable."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert r.stdout.count("bench ok") == 8 and "smoke ok" in r.stdout, tail


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_multi_rank_suites_and_bench_run_on_the_dry_run_backend():
    """world_size 2 (and 4) over gloo: the multi-GPU parity file (exchange, decomposed solvers, fused path) and bench.py
    launched exactly as the driver launches it for N > 1 (torch.distributed.run, one rank per "GPU"), with the dry-run
    backend's halo exchange / reductions travelling over the bootstrap group.  Checks the ranks' control flow (nobody
    enters a collective alone, rank 0 alone prints ONE line, both arms) -- not the NCCL path, which only GPUs can run."""
    env = dict(os.environ, CHMY_DRYRUN="1", CHMY_DRYRUN_NGPU="4", OMP_NUM_THREADS="2", CHMY_EXPERIMENTAL="1")   # + the gated peer-store cases
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "--runxfail", "-p", "no:cacheprovider", "-k",
                        "(2gpu and not peer and not stokes-40x33) or 4gpu-exchange-9x7x5 or 2gpu-exchange+peer-12x9",
                        "tests/test_z_b200_multigpu.py"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and " passed" in r.stdout, tail
    import json
    for extra in ("--n 24 20 16", "--workload diffusion2d --n 48 40"):
        # the driver's launch line for N > 1, with the dry-run main in bench.py's place
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dryrun_bench_main.py"), "--gpus", "2", "--steps", "6",
               "--warmup", "3"]
        r = subprocess.run(cmd, cwd=ROOT, env=dict(env, DRYRUN_BENCH_ARGS=extra), capture_output=True, text=True, timeout=600)
        tail = (r.stdout + r.stderr)[-3000:]
        assert r.returncode == 0, tail
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1, tail                                   # rank 0 alone prints, ONE line
        j = json.loads(lines[0])
        assert j["n_gpus"] == 2 and j["scaling"] == "weak" and j["value"] > 0 and j["config"]["proc_dims"][0] == 2
        assert "host_segment_error" not in j["e2e"] and j["cpu_baseline"] is None
    # the reference arm under torchrun: rank 0 alone works and prints
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--size", "96", "96", "96"]          # control flow of the ranks, not a measurement: a small grid
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1 and json.loads(lines[0])["impl"] == "reference", (r.stdout + r.stderr)[-2000:]
