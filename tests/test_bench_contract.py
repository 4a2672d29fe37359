"""CPU-only: bench.py's control flow and the JSON line it prints, checked without a GPU.

The B200 arm is driven against a MOCK of the `chmy_b200` host API (no arithmetic, no oracle: every device call is a
no-op that returns plausible numbers), so a typo in the measurement script cannot take the round-end bench down.
The reference arm (`--impl reference`) runs for real on a small grid: it times the CPU oracle and needs no GPU.
Numbers printed here mean nothing; the `-m gpu` suite and bench.py on the B200 box produce the measured ones."""
import importlib.util
import io
import json
import os
import subprocess
import sys
import types
from contextlib import redirect_stdout
from unittest import mock

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _F:
    """field stand-in: only what bench.py touches"""

    def __init__(self, dims):
        self.dims = tuple(dims)
        self.uploads = 0

    def from_host(self, a, lo, hi):
        self.uploads += 1


class _Tuple(list):
    pass


def _fake_api(nd, calls):
    from chmy_b200 import _lib as real_lib            # ctypes structs only: importable without a GPU
    ch = mock.MagicMock(name="chmy_b200")
    ch.ChmyError = RuntimeError
    ticks = {"launches": 0, "ev": 0}

    class Launch:
        blocking = False

        def __call__(self, *a, **k):
            ticks["launches"] += 1

    def mk_solver(arch, n, **kw):
        s = mock.MagicMock(name="solver")
        s.launch = Launch()
        s.grid = object()
        c = lambda extra=None: _F([x + (1 if extra == d else 0) for d, x in enumerate(n)])
        s.V = _Tuple(c(d) for d in range(nd))
        s.Pr, s.divV, s.C, s.T = c(), c(), c(), c()
        s.bc_V, s.bc_T = ((s.V[0], None),), ((s.T, None),)

        def step():
            ticks["launches"] += 2
        s.mechanics = s.step = step
        s.thermal = step
        return s

    drivers = types.SimpleNamespace(Stokes=mk_solver, Diffusion2D=mk_solver)
    ch.launch_count.side_effect = lambda arch: ticks["launches"]
    ch.fused_count.return_value = 7
    ch.event_elapsed_ms.return_value = 2.5
    ch.allreduce_max.side_effect = lambda arch, *v: tuple(float(x) for x in v)
    ch.maxabs.return_value = 0.25
    ch.pinned_array.side_effect = lambda arch, shape: np.zeros(tuple(shape), order="F")

    def interior(f, with_halo=False, out=None):
        calls["d2h"] += 1
        assert out is not None and out.shape == f.dims and out.flags.f_contiguous
        return out

    def set_(f, a):
        calls["h2d"] += 1
        assert a.shape == f.dims
    ch.interior.side_effect = interior
    ch.set_.side_effect = set_
    ch._lib = real_lib
    ch.drivers = drivers
    return ch, real_lib, drivers


@pytest.mark.parametrize("workload,n", [("stokes3d", (16, 12, 8)), ("stokes3d_thermal", (16, 12, 8)),
                                        ("stokes2d", (32, 24)), ("diffusion2d", (32, 24))])
def test_b200_arm_prints_the_contract_line(workload, n, monkeypatch):
    calls = {"h2d": 0, "d2h": 0}
    ch, real_lib, drivers = _fake_api(len(n), calls)
    monkeypatch.setitem(sys.modules, "chmy_b200", ch)
    monkeypatch.setitem(sys.modules, "chmy_b200._lib", real_lib)
    monkeypatch.setitem(sys.modules, "chmy_b200.drivers", drivers)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", workload, "--n", *map(str, n), "--steps", "4", "--warmup", "1"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_b200(bench.parse())
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1, lines                                    # ONE JSON line
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "roofline", "e2e", "cpu_baseline"):
        assert key in j, key
    assert j["metric"] == "T_eff" and j["unit"] == "GB/s" and j["dtype"] == "f64" and j["scaling"] == "weak"
    assert j["steps"] == 4 and j["warmup"] >= 3 and j["n_gpus"] == 1 and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    assert j["gpu_launches"] > 0
    r = j["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = j["e2e"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in e, key
    # the host-buffer segment ran: every state field went host -> device and device -> host inside the timed region
    nstate = 1 if workload == "diffusion2d" else len(n) + 1
    assert "host_segment_error" not in e, e
    assert calls["h2d"] == nstate and calls["d2h"] == 2 * nstate
    assert e["h2d_bytes_per_step"] > 8 * np.prod(n) * nstate / 4 and e["d2h_bytes_per_step"] > 8 * np.prod(n) * nstate / 4
    assert e["steady"]["h2d_bytes_per_step"] == 2 * __import__("ctypes").sizeof(real_lib.LaunchDesc)
    c = j["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in c, key
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0


def test_host_segment_failure_keeps_the_device_number(monkeypatch):
    calls = {"h2d": 0, "d2h": 0}
    ch, real_lib, drivers = _fake_api(3, calls)
    ch.pinned_array.side_effect = RuntimeError("cudaHostAlloc failed")          # -> pageable fallback

    def broken_interior(f, with_halo=False, out=None):
        raise MemoryError("no host memory")
    ch.interior.side_effect = broken_interior
    monkeypatch.setitem(sys.modules, "chmy_b200", ch)
    monkeypatch.setitem(sys.modules, "chmy_b200._lib", real_lib)
    monkeypatch.setitem(sys.modules, "chmy_b200.drivers", drivers)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "8", "8", "8", "--steps", "3", "--no-cpu-baseline"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_b200(bench.parse())
    j = json.loads(buf.getvalue().strip())
    assert j["value"] > 0 and "MemoryError" in j["e2e"]["host_segment_error"] and j["e2e"]["value"] > 0


def test_failure_inside_the_timed_host_segment_keeps_the_device_number(monkeypatch):
    """the upload of the segment's initial state (set!(f, A_host)) fails AFTER the buffers were prepared: the line still
    carries the device-timed value and the steady e2e, plus the reason"""
    calls = {"h2d": 0, "d2h": 0}
    ch, real_lib, drivers = _fake_api(3, calls)

    def broken_set(f, a):
        raise RuntimeError("chmy_b200 error -2: cudaMemcpy3D failed")
    ch.set_.side_effect = broken_set
    monkeypatch.setitem(sys.modules, "chmy_b200", ch)
    monkeypatch.setitem(sys.modules, "chmy_b200._lib", real_lib)
    monkeypatch.setitem(sys.modules, "chmy_b200.drivers", drivers)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "8", "8", "8", "--steps", "3", "--no-cpu-baseline"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_b200(bench.parse())
    j = json.loads(buf.getvalue().strip())
    assert j["value"] > 0 and "cudaMemcpy3D" in j["e2e"]["host_segment_error"] and j["e2e"]["value"] > 0
    assert j["roofline"]["traffic"] is None and j["roofline"]["traffic_other_size"]["n"] == "511x511x511"


def test_reference_arm_runs_on_the_cpu():
    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "24", "20", "12",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "T_eff" and j["unit"] == "GB/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_host_segment_is_refused_when_host_memory_is_short(monkeypatch):
    calls = {"h2d": 0, "d2h": 0}
    ch, real_lib, drivers = _fake_api(3, calls)
    monkeypatch.setitem(sys.modules, "chmy_b200", ch)
    monkeypatch.setitem(sys.modules, "chmy_b200._lib", real_lib)
    monkeypatch.setitem(sys.modules, "chmy_b200.drivers", drivers)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench = _load_bench()
    monkeypatch.setattr(bench, "host_mem_available", lambda: 1000)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "8", "8", "8", "--steps", "3", "--no-cpu-baseline"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_b200(bench.parse())
    j = json.loads(buf.getvalue().strip())
    assert "host memory" in j["e2e"]["host_segment_error"] and j["e2e"]["value"] > 0 and calls == {"h2d": 0, "d2h": 0}


@pytest.mark.parametrize("pd,n", [((2, 1, 1), (5, 4, 3)), ((2, 2, 1), (4, 3, 3)), ((2, 2, 2), (4, 3, 3)), ((4, 2), (5, 4)), ((1, 3), (4, 5))])
def test_bench_exchange_restatement_equals_the_oracle(pd, n):
    """bench.py's N-rank self-check may not touch the oracle at run time, so it restates the send / recv view algebra
    (communication_views.jl:1-34, exchange_halo.jl:73-84) in numpy; here that restatement is pinned on the oracle's
    lock-step world: index-encoded fields at four staggered locations, every process grid shape, corners included."""
    import oracle as o
    bench = _load_bench()
    nd, world = len(pd), int(np.prod(pd))
    topos = [o.Topology(world, pd, r) for r in range(world)]
    n_g = tuple(a * p for a, p in zip(n, pd))
    ogs = [o.local_grid((-1.0,) * nd, (2.0,) * nd, n_g, t) for t in topos]
    locs = [(0,) * nd, (1,) + (0,) * (nd - 1), (0,) * (nd - 1) + (1,), (1,) * nd]
    ofs = [[o.Field(og, l) for l in locs] for og in ogs]
    parents = []
    for r in range(world):
        for f in ofs[r]:
            f.data[...] = r * 1.0e6 + np.arange(f.data.size, dtype=np.float64).reshape(f.sdims, order="F")
        parents.append([f.data.copy() for f in ofs[r]])
    o.bc_world(ogs, [o.batch(ogs[r], exchange=tuple(ofs[r])) for r in range(world)], topos)
    want = bench._exchange_expected(parents, locs, pd, [t.coords for t in topos], n)
    for r in range(world):
        for q in range(len(locs)):
            assert np.array_equal(want[r][q], ofs[r][q].data), (r, locs[q])
