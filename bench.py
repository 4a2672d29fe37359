#!/usr/bin/env python
"""
bench.py -- T_eff of the pseudo-transient solvers' hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...] [--n ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one PT iteration of the named workload through the reference-facing API (`launch(arch, grid, op => args;
bc=batch(...))`, examples/stokes_3d_inc_ve_T_mpi_perf.jl:183-187): update_stress! + update_velocity! + the velocity
boundary batch / halo exchange.  T_eff = nIO * 8 B * prod(n_local) / t_it with the reference's nIO
(37 for 3D Stokes mechanics, stokes_3d_inc_ve_T_mpi_perf.jl:210; 7 for 2D diffusion, diffusion_2d_perf.jl:65).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference itself is Julia and cannot
run in this image) on a bounded slab of the same workload with all host threads.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (default local size, nIO, description)
    "stokes3d": ((767, 767, 767), 37, "examples/stokes_3d_inc_ve_T_mpi_perf.jl 3D Stokes PT (mechanics), Float64, 767^3 per GPU"),
    "stokes2d": ((8191, 8191), 22, "examples/stokes_2d_inc_ve_T.jl 2D Stokes PT (mechanics), Float64, 8191^2"),
    "diffusion2d": ((16383, 16383), 7, "examples/diffusion_2d_perf.jl 2D diffusion, Float64, 16383^2"),
    # outer steps it > 1: mechanics + thermal sub-step (nIO 37 + 9, stokes_3d_inc_ve_T_mpi.jl:206; 2D by the same rule)
    "stokes3d_thermal": ((767, 767, 767), 46, "examples/stokes_3d_inc_ve_T.jl 3D Stokes PT + thermal sub-step, Float64, 767^3"),
    "stokes2d_thermal": ((8191, 8191), 29, "examples/stokes_2d_inc_ve_T.jl 2D Stokes PT + thermal sub-step, Float64, 8191^2"),
}
# real (perfect-reuse) array passes of the dominant kernel, per cell of its (n+2)^N launch range
DOMINANT = {"stokes3d": ("update_stress!", 24), "stokes2d": ("update_stress!", 14), "diffusion2d": ("update_C!", 4),
            "stokes3d_thermal": ("update_stress!", 24), "stokes2d_thermal": ("update_stress!", 14)}


def measured_traffic(workload, n, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[f"{workload}:{'x'.join(map(str, n))}:{kernel}"]["bytes"]
    except Exception:
        return None


def nearest_traffic(workload, kernel):
    """When no capture exists at the benchmarked size: the committed capture of the same kernel at another size, with its
    ratio to the algorithmic bytes of THAT size (evidence for how close the kernel's DRAM traffic is to its floor)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for key, v in t.items():
            parts = key.split(":", 2)
            if len(parts) == 3 and parts[0] == workload and parts[2] == kernel and "algorithmic" in v:
                return {"n": parts[1], "bytes": v["bytes"], "algorithmic": v["algorithmic"],
                        "ratio": v["bytes"] / v["algorithmic"], "source": v.get("source")}
    except Exception:
        pass
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="stokes3d", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--size", type=int, nargs="*", default=None, help="override the local grid size (parity/debug runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the untimed multi-rank correctness check")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the other BASELINE configs (extra.workloads)")
    ap.add_argument("--split", default="on", choices=["auto", "on", "off", "always"],
                    help="launches with boundary batches -- on (= auto, the default): the batches and the halo exchange overlap "
                         "the kernel (fused 3D sweep: boundary tiles first + a retire counter, one launch; plain kernels: the "
                         "reference's inner region + slabs); off: one kernel, then batches + exchange on one stream "
                         "(results are identical)")
    ap.add_argument("--no-split", action="store_true", help="same as --split off")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="multi-GPU: transport of the halo exchange -- nccl = pack + ncclSend/ncclRecv + unpack (the measured path), "
                         "peer = EXPERIMENTAL: the pack kernels store into the neighbour's HBM over NVLink, sequence flags instead "
                         "of the NCCL group (chmy_set_exchange_mode; results are identical)")
    ap.add_argument("--fused", type=int, default=3, choices=[0, 1, 3],
                    help="lazily fuse pairs of launches into one sweep (chmy_set_fusion): 3 (default) = every pair with a sweep "
                         "(3D stress + velocity; 2D stress + velocity, compute_q + update_C, thermal flux + update in 2D / 3D); "
                         "1 = only the 3D stress + velocity sweep; 0 = the two tuned kernels")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


def host_mem_available():
    """bytes of host memory this process may still take: the smaller of the system's available memory and the cgroup's
    remaining allowance (a container limit is what an out-of-memory kill enforces).  None if nothing can be read."""
    vals = []
    try:
        import psutil
        vals.append(int(psutil.virtual_memory().available))
    except Exception:
        pass
    for lim, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            l = open(lim).read().strip()
            if l != "max" and int(l) < (1 << 60):
                vals.append(int(l) - int(open(cur).read().strip()))
        except Exception:
            pass
    return min(vals) if vals else None


def a_eff_bytes(workload, n_local):
    return WORKLOADS[workload][1] * 8.0 * float(math.prod(n_local))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(workload, n_full, steps, warmup, *, full, budget_s):
    """The oracle (C restatement of the reference's kernels, OpenMP) on the host's cores: the stand-in for the reference's
    KernelAbstractions-CPU path, which cannot run here (no Julia).  Built on this host with -O3 -march=native
    (-ffp-contract=off: same bits), ALL cores even when a launcher exported OMP_NUM_THREADS=1 for its workers.
    full: the whole grid when host memory allows (the reference arm), else a bounded slab of it (the in-line baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["CHMY_ORACLE_NATIVE"] = "1"
    os.environ["OMP_NUM_THREADS"] = str(cores)            # read by libgomp when it loads (torchrun sets it to 1)
    # worker threads spin between the (many, short) parallel regions of a step instead of sleeping: on virtualised hosts a
    # sleeping team can take milliseconds to wake up, which would be charged to the CPU arm
    os.environ.setdefault("OMP_WAIT_POLICY", "active")
    os.environ.setdefault("OMP_PROC_BIND", "close")
    import numpy as np
    import oracle as o
    import drivers as OD
    cores = o.set_num_threads(cores)
    os.environ.pop("CHMY_ORACLE_NATIVE", None)            # the binding is loaded: leave the environment as it was found
    nd = len(n_full)
    nfields = 3 if workload == "diffusion2d" else (25 if nd == 3 else 19)
    avail = host_mem_available()

    def fits(n):
        need = nfields * 8.0 * float(math.prod(x + 4 for x in n))
        return avail is None or 1.4 * need < avail

    n = tuple(n_full)
    if not full:          # bounded slab: a quarter of the 3D grid along z, 4096 rows of a 2D grid
        n = tuple(n_full[:-1]) + (min(n_full[-1], 192 if nd == 3 else 4096),)
    while not fits(n) and n[-1] > 16:
        n = tuple(n[:-1]) + (n[-1] // 2,)
    if workload == "diffusion2d":
        sol = OD.Diffusion2D(n, outer_width=(128, 8), C0=np.random.default_rng(0).random(n))
        step = sol.step
    else:
        sol = OD.Stokes(n, re_m=2.5 * math.pi, rho_g_function=True, adv_coef=0.01)
        sol.begin_time_step()
        if workload.endswith("_thermal"):
            def step():
                sol.mechanics()
                sol.thermal()
        else:
            step = sol.mechanics
    whole = n == tuple(n_full)
    sample = ("the whole " + "x".join(map(str, n)) + " grid") if whole else \
        ("x".join(map(str, n)) + " slab of the " + "x".join(map(str, n_full)) + " grid")
    t0 = time.perf_counter()
    wdone = 0
    while wdone < max(1, warmup) and (wdone < 1 or time.perf_counter() - t0 < 0.25 * budget_s):
        step()
        wdone += 1
    t0 = time.perf_counter()
    done = 0
    while done < steps and (done < 2 or time.perf_counter() - t0 < budget_s):
        step()
        done += 1
    dt = (time.perf_counter() - t0) / done
    teff = a_eff_bytes(workload, n) / dt / 1e9
    return {"value": teff, "unit": "GB/s", "cores": cores, "kind": "port", "n_sample": list(n), "same_config": whole,
            "steps": done, "warmup": wdone,
            "sample": f"{sample}; {wdone} warm-up + {done} timed steps, {dt*1e3:.1f} ms/step; oracle/chmy_oracle.c built on this host "
                      f"(gcc -O3 -march=native -fopenmp -ffp-contract=off), {cores} threads; the Julia reference cannot run in this image"}, dt, done, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_full = tuple(args.n) if args.n else WORKLOADS[args.workload][0]
    cb, dt, done, n = cpu_arm(args.workload, n_full, args.steps, args.warmup, full=True, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "T_eff", "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": done, "warmup": cb["warmup"], "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][2], "n_local": list(n), "sample": cb["sample"], "nIO": WORKLOADS[args.workload][1],
                   "note": "one host process with all cores whatever --gpus says: the CPU stand-in does not shard"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the other BASELINE configs
def extra_workloads(steps=50, warmup=10):
    """BASELINE.json configs 2-4 next to the headline (config 5): each in its own process (the headline alone holds ~130 GB of
    fields), device-timed like the headline, two-kernel and fused.  Returns the condensed lines."""
    out = []
    for wl, fusions in (("diffusion2d", (0, 3)), ("stokes2d", (0, 3)), ("stokes2d_thermal", (0, 3)), ("stokes3d_thermal", (1, 3))):
        for fu in fusions:
            cmd = [sys.executable, os.path.abspath(__file__), "--workload", wl, "--fused", str(fu), "--steps", str(steps), "--warmup", str(warmup),
                   "--no-e2e", "--no-cpu-baseline", "--no-extra"]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
                d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                out.append({"workload": d["config"]["workload"], "n": d["config"]["n_local"], "nIO": d["config"]["nIO"], "fused": fu,
                            "steps": d["steps"], "warmup": d["warmup"], "ms_per_step": d["ms_per_step"], "T_eff": d["T_eff_per_gpu"],
                            "frac_of_hbm_peak": d["frac_of_hbm_peak"], "gpu_launches": d["gpu_launches"],
                            "roofline": {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "kernel_ms", "step_kernels_ms")}})
            except Exception as ex:
                out.append({"workload": wl, "fused": fu, "error": f"{type(ex).__name__}: {ex}"})
    return out


# ------------------------------------------------------------------------------------------------ multi-GPU self-check
def _exchange_expected(parents, locs, dims, coords_of, n):
    """communication_views.jl:1-34 + exchange_halo.jl:73-84 restated on host copies of the padded arrays: for D = N..1,
    every rank's recv slab (logical 0 | d+1 along D, whole padded transverse extent) := the neighbour's send slab
    (logical 1+overlap | d-overlap).  parents[r][q]: padded array of field q on rank r (array index = logical + 1)."""
    import numpy as np
    nd = len(dims)
    rank_of = {tuple(c): r for r, c in enumerate(coords_of)}
    out = [[a.copy() for a in fs] for fs in parents]
    for D in range(nd - 1, -1, -1):
        snap = [[a.copy() for a in fs] for fs in out]      # both sides of a dim move concurrently, dims sequentially
        for r, c in enumerate(coords_of):
            for side, delta in ((0, -1), (1, +1)):
                cc = list(c); cc[D] += delta
                if not (0 <= cc[D] < dims[D]):
                    continue
                nb = rank_of[tuple(cc)]
                for q, loc in enumerate(locs):
                    ov = 1 if loc[D] else 0
                    d = n[D] + ov
                    recv = 1 if side == 0 else d + 2
                    send = (d - ov + 1) if side == 0 else (2 + ov)     # the neighbour's opposite side
                    sl_r = [slice(None)] * nd; sl_r[D] = recv
                    sl_s = [slice(None)] * nd; sl_s[D] = send
                    out[r][q][tuple(sl_r)] = snap[nb][q][tuple(sl_s)]
    return out


def multi_gpu_check(ch, BD, arch, backend, world, rank, local_rank, fused):
    """Untimed correctness check of the N-rank path through the product API only (no oracle): (1) exchange_halo! on
    index-encoded fields (rank*1e6 + linear storage index) must equal the send/recv view algebra bit for bit, corners
    included; (2) the N-rank decomposed Stokes solve of a small global grid must agree with the same global grid solved on
    ONE GPU (rank 0, a second single-device context) to <= 1e-12 relative, fields and residual history (SURVEY 8c:
    not bitwise, the sub-axis spacing is recomputed per rank, distributed_grid.jl:19-36)."""
    import numpy as np
    topo = arch.topology
    pd, comm = topo.dims, topo.comm
    nd = len(pd)
    out = {"world": world, "proc_dims": list(pd)}
    # ---- (1) exchange_halo!, examples/exchange_halo.jl:20-30 with index-encoded contents
    n = (9, 7, 5)[:nd]
    n_g = tuple(a * p for a, p in zip(n, pd))
    g = ch.UniformGrid(arch, origin=(-1.0,) * nd, extent=(2.0,) * nd, dims=n_g)
    locs = [(0,) * nd, (1,) + (0,) * (nd - 1), (0,) * (nd - 1) + (1,), (1,) * nd]
    fs = [ch.Field(arch, g, tuple(ch.Vertex() if x else ch.Center() for x in l)) for l in locs]
    mine = []
    for f, l in zip(fs, locs):
        shape = tuple(a + x + 4 for a, x in zip(n, l))
        a = rank * 1.0e6 + np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape, order="F")
        f.from_host(a.copy(), [-1] * nd, [d + 2 for d in f.dims])
        mine.append(a)
    allp = comm.allgather_obj((tuple(topo.cart_coords), mine))
    ch.exchange_halo_(arch, g, *fs)
    want = _exchange_expected([p for _, p in allp], locs, pd, [c for c, _ in allp], n)[rank]
    bad = 0
    for f, w in zip(fs, want):
        bad += int((f.parent() != w).sum())
    for f in fs:
        f.free()
    (bad_all,) = ch.allreduce_max(arch, float(bad))
    out["exchange_bit_exact"] = bad_all == 0.0
    out["exchange_cells_wrong_max_over_ranks"] = int(bad_all)
    # ---- (2) decomposed solve vs one GPU
    nl = (24, 20, 16)[:nd]
    kw = dict(re_m=2.5 * math.pi, rho_g_function=True, adv_coef=0.01, blocking=False)
    ch.set_fusion(arch, int(bool(fused)))
    sol = BD.Stokes(arch, nl, outer_width=(8, 4, 3)[:nd], **kw)
    hist = sol.run(2, 20, 10)
    blocks = {k: np.asarray(ch.interior(f)) for k, f in sol.fields().items()}
    allb = comm.allgather_obj((tuple(topo.cart_coords), blocks))
    worst, worst_name, hist_rel = 0.0, None, 0.0
    if rank == 0:
        arch1 = ch.Arch(backend, device_id=local_rank + 1)
        ch.set_fusion(arch1, int(bool(fused)))
        ng = tuple(a * p for a, p in zip(nl, pd))
        ref = BD.Stokes(arch1, ng, outer_width=(8, 4, 3)[:nd], **kw)
        href = ref.run(2, 20, 10)
        assert len(href) == len(hist), (len(href), len(hist))
        for a, b in zip(hist, href):
            assert a[:2] == b[:2]
            for x, y in zip(a[2:], b[2:]):
                hist_rel = max(hist_rel, abs(x - y) / max(abs(y), 1e-300))
        for k, f in ref.fields().items():
            G = np.asarray(ch.interior(f))
            den = max(float(np.abs(G).max()), 1e-300)
            for c, blk in allb:
                b = blk[k]
                sl = tuple(slice(ci * ni, ci * ni + sz) for ci, ni, sz in zip(c, nl, b.shape))
                e = float(np.abs(G[sl] - b).max()) / den
                if e > worst:
                    worst, worst_name = e, f"{k} @ coords {c}"
        arch1.close()
    (worst, hist_rel) = ch.allreduce_max(arch, worst, hist_rel)
    out.update(max_rel=worst, history_max_rel=hist_rel, worst=worst_name, n_local=list(nl), pt_iterations=40,
               tolerance=1e-12, ok=bool(bad_all == 0.0 and worst <= 1e-12 and hist_rel <= 1e-12),
               what="N-rank decomposed 3D Stokes (2 outer steps x 20 PT iterations, thermal on in step 2, fused sweep as timed) vs the "
                    "same global grid on one GPU; exchange_halo! on index-encoded fields vs the send/recv view algebra")
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    import numpy as np
    import ctypes as C
    import chmy_b200 as ch
    from chmy_b200 import _lib as L
    from chmy_b200 import drivers as BD

    backend = ch.B200Backend()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)   # bootstrap only (NCCL id broadcast)
        nd = len(WORKLOADS[args.workload][0])
        arch = ch.Arch(backend, ch.TorchDistComm(), (0,) * nd, device_id=local_rank + 1)
        pdims = arch.topology.dims
        ch.set_exchange_mode(arch, args.exchange)
    else:
        arch = ch.Arch(backend, device_id=local_rank + 1)
        pdims = None

    wl = args.workload
    n = tuple(args.n) if args.n else WORKLOADS[wl][0]
    split_mode = "off" if args.no_split else args.split
    ch.set_launch_split(arch, "always" if split_mode == "always" else split_mode != "off")
    fused2d = args.fused == 3 and not wl.startswith("stokes3d")      # 2D flux -> update sweeps (ops_fused2d.cu)
    fused_t3 = args.fused == 3 and wl == "stokes3d_thermal"          # 3D thermal sweep (fused_thermal3.cuh)
    fused = (bool(args.fused) and wl.startswith("stokes3d")) or fused2d
    ch.set_fusion(arch, 3 if (fused2d or fused_t3) else int(fused))
    if wl == "diffusion2d":
        sol = BD.Diffusion2D(arch, n, outer_width=(128, 8), C0=None, blocking=False)
        # uniform [0,1) initial condition generated on the host in strips (the reference uses rand())
        rng = np.random.default_rng(rank)
        strip = 1024
        for j0 in range(1, n[1] + 1, strip):
            j1 = min(n[1], j0 + strip - 1)
            sol.C.from_host(rng.random((n[0], j1 - j0 + 1)), [1, j0], [n[0], j1])
        ch.bc_(arch, sol.grid, (sol.C, ch.Neumann()), exchange=sol.C)
        step = sol.step
        sub = [("compute_q!", lambda: sol.launch(arch, sol.grid, (ch.compute_q_, (sol.q, sol.C, sol.chi, sol.grid)))),
               ("update_C!", lambda: sol.launch(arch, sol.grid, (ch.update_C_, (sol.C, sol.q, sol.dt, sol.grid)),
                                                bc=ch.batch(sol.grid, (sol.C, ch.Neumann()), exchange=sol.C)))]
        if fused:
            q0, q1 = sub[0][1], sub[1][1]
            sub = [("compute_q!+update_C! (fused sweep)", lambda: (q0(), q1()))]
        metric_field = sol.C
    else:
        sol = BD.Stokes(arch, n, re_m=2.5 * math.pi, rho_g_function=True,
                        outer_width=(128, 8, 4) if len(n) == 3 else (128, 8), adv_coef=0.01, blocking=False)
        sol.begin_time_step()
        g = sol.grid
        if wl.endswith("_thermal"):
            def step():
                sol.mechanics()
                sol.thermal()
        else:
            step = sol.mechanics
        sub = [("update_stress!", lambda: sol.launch(arch, g, (ch.update_stress_, (sol.tau, sol.Pr, sol.divV, sol.V, sol.tau_old,
                                                     sol.eta, sol.eta_ve, sol.G, sol.dt, sol.dtau_Pr, sol.dtau_r, g)))),
               ("update_velocity!", lambda: sol.launch(arch, g, (ch.update_velocity_, (sol.V, sol.r_V, sol.Pr, sol.tau, sol.rho_g,
                                                       sol.eta_ve, sol.nudtau, g)), bc=ch.batch(g, *sol.bc_V, exchange=sol.exch_V)))]
        if fused:    # the two launches run as one sweep: time them together (an event between them would un-fuse them)
            s0, s1 = sub[0][1], sub[1][1]
            sub = [("update_stress!+update_velocity! (fused sweep)", lambda: (s0(), s1()))]
        if wl.endswith("_thermal"):
            sub += [("update_thermal_flux!", lambda: sol.launch(arch, g, (ch.update_thermal_flux_, (sol.qT, sol.T, sol.V, sol.lam, g)))),
                    ("update_thermal!", lambda: sol.launch(arch, g, (ch.update_thermal_, (sol.T, sol.T_old, sol.qT, sol.dt, g)),
                                                           bc=ch.batch(g, *sol.bc_T, exchange=sol.T)))]
            if fused2d or fused_t3:
                t0_, t1_ = sub[-2][1], sub[-1][1]
                sub = sub[:-2] + [("update_thermal_flux!+update_thermal! (fused sweep)", lambda: (t0_(), t1_()))]
        metric_field = sol.divV
    ch.synchronize(arch)

    K, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        step()
    ch.synchronize(arch)
    ch.barrier(arch)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ch.launch_count(arch)
    ch.event_record(arch, 0)
    for _ in range(K):
        step()
    ch.event_record(arch, 1)
    ch.synchronize(arch)
    ms_local = ch.event_elapsed_ms(arch, 0, 1)
    l1 = ch.launch_count(arch)
    nfused = ch.fused_count(arch)
    noverl = ch.overlapped_count(arch)
    nfallback = ch.fusion_fallback_count(arch)
    # how the fused 3D sweep divided by its uniform scalars (0: four operations, 1: div.rn.f64, 2: two operations, proven exact)
    divmode = {0: "4 operations (Markstein)", 1: "div.rn.f64", 2: "2 operations (proven exact for these divisors)"}.get(ch.last_division_mode(arch)) \
        if (fused and wl.startswith("stokes3d")) else None
    xstats = dict(zip(("peer", "nccl"), ch.exchange_stats(arch))) if world > 1 else None     # rank 0's messages so far
    clocks = sampler.stop() if rank == 0 else None
    ch.barrier(arch)
    (ms_max,) = ch.allreduce_max(arch, ms_local) if world > 1 else (ms_local,)
    # every rank's own device time of the K steps (a slow GPU or a late rank shows up here; the line reports the max)
    per_rank = list(ch.allreduce_max(arch, *[ms_local / K if r == rank else 0.0 for r in range(world)])) if world > 1 else [ms_local / K]
    t_it = ms_max / K * 1e-3
    teff_gpu = a_eff_bytes(wl, n) / t_it / 1e9

    # ---- per-kernel device time of the two launches of a step (events on the launching stream)
    KK = min(K, 20)
    slot = 10
    for _ in range(KK):
        for _, fn in sub:
            ch.event_record(arch, slot); slot += 1
            fn()
        ch.event_record(arch, slot); slot += 1
    ch.synchronize(arch)
    per = {nm: 0.0 for nm, _ in sub}
    s = 10
    for _ in range(KK):
        for nm, _ in sub:
            per[nm] += ch.event_elapsed_ms(arch, s, s + 1)
            s += 1
        s += 1
    per = {k: v / KK for k, v in per.items()}
    # the fused 3D sweep and its boundary batch are ONE launch call but two kernels: events recorded inside the library right
    # before / after the sweep kernel (chmy_time_fused_sweep) isolate the sweep; the rest of the call is the batch kernel
    if fused and wl.startswith("stokes3d") and len(sub) >= 1:
        nm0 = sub[0][0]
        ch.time_fused_sweep(arch, 8, 9)
        tk = 0.0
        for _ in range(KK):
            sub[0][1]()
            ch.synchronize(arch)
            tk += ch.event_elapsed_ms(arch, 8, 9)
        ch.time_fused_sweep(arch, -1, -1)
        call_ms = per[nm0]
        per[nm0] = tk / KK
        per["boundary batches of that launch (k_bc_all) + launch gap"] = max(call_ms - tk / KK, 0.0)
    dom_name, dom_passes = DOMINANT[wl]
    if fused:        # 3D: R16 + W14 array passes; 2D Stokes: R9 + W9; diffusion: R1 + W3
        dom_name, dom_passes = sub[0][0], (30 if len(n) == 3 else (4 if wl == "diffusion2d" else 18))
    cells_launch = float(math.prod(x + 2 for x in n))
    alg_bytes = dom_passes * 8.0 * cells_launch
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (per[dom_name] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(wl, n, dom_name), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": per[dom_name],
                "step_kernels_ms": per, "share_of_step": per[dom_name] / sum(per.values())}
    if roofline["traffic"] is None:
        roofline["traffic_other_size"] = nearest_traffic(wl, dom_name)

    # ---- end to end through the public API (chmy_b200.Launcher / set! / interior) with HOST buffers.
    #      Timed with the host clock: a K-iteration solve segment whose primary unknowns start and end in host memory --
    #        set!(f, A_host) for every state field (H2D from pinned host memory, field.jl:98),
    #        K PT iterations with the reference's blocking launches (KernelLaunch.jl:117) + one max|residual| read-back each,
    #        Array(interior(f)) of every state field into pinned host memory (D2H, field.jl:33-37).
    #      `steady` is the same loop without the state transfers (fields resident, as in the reference's own perf driver).
    e2e = None
    if not args.no_e2e:
        state = [sol.C] if wl == "diffusion2d" else list(sol.V) + [sol.Pr]
        sol.launch.blocking = True
        for _ in range(2):
            step(); ch.maxabs(metric_field)
        ch.synchronize(arch); ch.barrier(arch)
        t0 = time.perf_counter()
        for _ in range(K):
            step()
            ch.maxabs(metric_field)
        ch.synchronize(arch)
        dt_steady = time.perf_counter() - t0
        (dt_steady,) = ch.allreduce_max(arch, dt_steady) if world > 1 else (dt_steady,)
        desc_bytes = 2 * C.sizeof(L.LaunchDesc)
        e2e = {"value": world * a_eff_bytes(wl, n) / (dt_steady / K) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": desc_bytes, "d2h_bytes_per_step": 8, "ms_per_step": dt_steady / K * 1e3,
               "what": "blocking launches (KernelLaunch.jl:117 semantics) through chmy_b200.Launcher + one max|residual| "
                       "read-back per step; fields resident in HBM (host-buffer segment failed, see `host_segment_error`)"}
        views, host_kind, err = None, None, None
        try:
            # every rank of the node holds its own copy of the state in host memory: refuse rather than push the box into
            # an out-of-memory kill
            need = world * sum(8 * int(np.prod(f.dims, dtype=np.int64)) for f in state)
            avail = host_mem_available()
            if avail is not None and avail < 1.5 * need:
                raise MemoryError(f"host-buffer segment needs {need / 1e9:.1f} GB of host memory on this node, {avail / 1e9:.1f} GB available")
            try:
                views, host_kind = [ch.pinned_array(arch, f.dims) for f in state], "pinned (chmy_host_alloc)"
            except ch.ChmyError:
                views, host_kind = [np.empty(tuple(f.dims), dtype=np.float64, order="F") for f in state], \
                    "pageable (pinned allocation failed)"
            # the segment's initial state in host memory: what the fields hold now (read back untimed)
            for f, v in zip(state, views):
                ch.interior(f, out=v)
        except Exception as ex:      # the host-buffer segment must never take the device-timed number down
            err = f"{type(ex).__name__}: {ex}"
        (bad,) = ch.allreduce_max(arch, float(err is not None)) if world > 1 else (float(err is not None),)
        if bad:                      # decided by all ranks together: nobody enters the collective timing alone
            e2e["host_segment_error"] = err or "failed on another rank"
        else:
            seg_err, dt_e2e = None, 0.0
            try:
                ch.synchronize(arch); ch.barrier(arch)
                t0 = time.perf_counter()
                for f, v in zip(state, views):
                    ch.set_(f, v)                                     # H2D of the segment's initial state
                t1 = time.perf_counter()
                for _ in range(K):
                    step()
                    ch.maxabs(metric_field)
                ch.synchronize(arch)
                t2 = time.perf_counter()
                for f, v in zip(state, views):
                    ch.interior(f, out=v)                             # D2H of the segment's result
                t3 = time.perf_counter()
                dt_e2e = t3 - t0
            except Exception as ex:  # same rule: report, keep the device-timed number and the steady e2e
                seg_err = f"{type(ex).__name__}: {ex}"
            (dt_e2e, bad) = ch.allreduce_max(arch, dt_e2e, float(seg_err is not None)) if world > 1 else (dt_e2e, float(seg_err is not None))
            if bad:
                e2e["host_segment_error"] = seg_err or "failed on another rank"
            else:
                fbytes = [8 * int(np.prod(f.dims, dtype=np.int64)) for f in state]
                h2d = d2h = sum(fbytes)
                e2e = {"value": world * a_eff_bytes(wl, n) * K / dt_e2e / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": h2d / K + desc_bytes, "d2h_bytes_per_step": d2h / K + 8,
                       "ms_per_step": dt_e2e / K * 1e3, "segment_steps": K, "host_memory": host_kind,
                       "upload_ms": (t1 - t0) * 1e3, "iterate_ms": (t2 - t1) * 1e3, "download_ms": (t3 - t2) * 1e3,
                       "steady": {"value": world * a_eff_bytes(wl, n) / (dt_steady / K) / 1e9, "ms_per_step": dt_steady / K * 1e3,
                                  "h2d_bytes_per_step": desc_bytes, "d2h_bytes_per_step": 8},
                       "what": f"K={K}-iteration solve segment through chmy_b200 (set!(f, A_host) of {len(state)} state fields from host "
                               "memory -> K x [blocking launches (KernelLaunch.jl:117) + one max|residual| read-back] -> "
                               "Array(interior(f)) of the state fields into host memory), host clock, max over ranks; "
                               "`steady` = the same loop with the fields resident, as the reference's perf driver times it"}
        del views
        sol.launch.blocking = False

    # ---- N > 1: the halo exchange of one iteration on its own (pack -> transport -> unpack, z then y then x), device-timed
    xchg = None
    if world > 1 and wl.startswith("stokes"):
        try:
            ch.synchronize(arch); ch.barrier(arch)
            for _ in range(3):
                ch.exchange_halo_(arch, sol.grid, *sol.V, blocking=False)
            ch.synchronize(arch); ch.barrier(arch)
            ch.event_record(arch, 2)
            for _ in range(20):
                ch.exchange_halo_(arch, sol.grid, *sol.V, blocking=False)
            ch.event_record(arch, 3)
            ch.synchronize(arch)
            (xms,) = ch.allreduce_max(arch, ch.event_elapsed_ms(arch, 2, 3) / 20)
            per = {}
            for D in range(len(n)):     # one dimension at a time: both sides of dim D through an exchange-only bc! batch (a collective every rank joins)
                if pdims[D] == 1:
                    continue
                spec = {"xyz"[D]: tuple(sol.V)}
                ch.bc_(arch, sol.grid, exchange=spec, blocking=False)
                ch.synchronize(arch); ch.barrier(arch)
                ch.event_record(arch, 4)
                for _ in range(10):
                    ch.bc_(arch, sol.grid, exchange=spec, blocking=False)
                ch.event_record(arch, 5)
                ch.synchronize(arch)
                (t,) = ch.allreduce_max(arch, ch.event_elapsed_ms(arch, 4, 5) / 10)
                per[f"dim{D + 1}"] = t
            face = [8.0 * sum(int(np.prod([f.dims[a] + 4 for a in range(len(n)) if a != D])) for f in sol.V) for D in range(len(n))]
            xchg = {"ms_per_exchange_alone": xms, "ms_per_dim_alone": per, "connected_dims": [D + 1 for D in range(len(n)) if pdims[D] > 1],
                    "bytes_per_side_per_dim": face,
                    "what": "exchange_halo!(arch, grid, V...) alone on an idle device, max over ranks: what one iteration's exchange "
                            "costs when nothing hides it (the timed loop overlaps it with the sweep unless --split off)"}
        except Exception as ex:
            xchg = {"error": f"{type(ex).__name__}: {ex}"}

    mgc = None
    if world > 1 and not args.no_check:
        try:
            mgc = multi_gpu_check(ch, BD, arch, backend, world, rank, local_rank, fused and wl.startswith("stokes3d"))
        except Exception as ex:
            mgc = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_arm(wl, n, 8, 2, full=False, budget_s=15.0)[0]
        except Exception as e:      # the CPU arm must never take the GPU number down
            cb = {"value": None, "unit": "GB/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": "T_eff", "value": teff_gpu * world, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_it * 1e3, "ms_per_step_by_rank": per_rank, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][2], "n_local": list(n), "proc_dims": list(pdims) if pdims else [1] * len(n),
                       "nIO": WORKLOADS[wl][1], "fused_sweep": fused, "split_launches": split_mode, "exchange": (args.exchange if world > 1 else None), "A_eff_GB_per_gpu": a_eff_bytes(wl, n) / 1e9,
                       "l2": "inputs larger than L2 (every field >= 2 GB; 126 MB L2), no flush needed",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "T_eff_per_gpu": teff_gpu, "frac_of_hbm_peak": teff_gpu / peak, "hbm_peak": peak, "hbm_peak_source": peak_src,
            "clocks": clocks, "gpu_launches": int(l1 - l0), "launches_per_step": (l1 - l0) / K, "fused_sweeps": int(nfused), "fusion_fallbacks": int(nfallback), "division": divmode, "overlapped_launches": int(noverl), "exchange_msgs": xstats, "roofline": roofline, "e2e": e2e, "cpu_baseline": cb,
            "multi_gpu_check": mgc, "exchange_alone": xchg,
        }
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    # release the device (the headline alone holds ~130 GB of fields and shadows) before the other configs run in their own processes
    metric_field = state = sub = step = None
    for f in sol.fields().values():
        f.free()
    del sol
    arch.close()
    if rank == 0:
        if world == 1 and wl == "stokes3d" and not args.no_extra and not args.n:
            import gc
            gc.collect()
            line["extra"] = {"workloads": extra_workloads(),
                             "what": "BASELINE.json configs 2-4 (+ the 2D thermal sub-step), each in its own process: device-timed "
                                     "T_eff like the headline (10 warm-up + 50 steps), two tuned kernels (fused 0 | 1) vs the fused sweeps (3)"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
