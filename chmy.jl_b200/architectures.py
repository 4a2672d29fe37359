"""
Architectures: device + stream ownership behind the C ABI.
Mirrors src/Architectures.jl:14-89, ext/ChmyCUDAExt/ChmyCUDAExt.jl:7-23 and
src/Distributed/distributed_architecture.jl:6-75.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L


class B200Backend:
    """The KernelAbstractions-style backend tag of this path (stands where `CUDABackend()` stands in the
    reference's drivers).  There is exactly one backend: hand-written CUDA for sm_100a."""

    def __repr__(self):
        return "B200Backend()"


class Architecture:
    pass


class SingleDeviceArchitecture(Architecture):
    """SingleDeviceArchitecture{B,D} (Architectures.jl:21-28): owns the device context and its streams."""

    def __init__(self, backend, device_id: int = 1):
        if isinstance(backend, SingleDeviceArchitecture):      # SingleDeviceArchitecture(arch::Architecture), Architectures.jl:30
            backend, device_id = backend.backend, backend.device_id
        if not isinstance(backend, B200Backend):
            raise TypeError("this path has a single backend: B200Backend()")
        self.backend = backend
        self.device_id = int(device_id)
        h = C.c_void_p()
        L.check(L.lib().chmy_ctx_create(self.device_id, C.byref(h)))   # set_device! + streams
        self._ctx = h

    @property
    def ctx(self):
        if self._ctx is None:
            raise L.ChmyError("architecture was destroyed")
        return self._ctx

    @property
    def device(self):
        """arch.device (Architectures.jl:23): the reference's 1-based device id"""
        return self.device_id

    def close(self):
        if getattr(self, "_ctx", None) is not None:
            L.lib().chmy_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistributedArchitecture(Architecture):
    """DistributedArchitecture (distributed_architecture.jl:6-34): child single-device arch + CartesianTopology."""

    def __init__(self, child_arch: SingleDeviceArchitecture, topology, gpu_aware: bool = True):
        self.child_arch = child_arch
        self.topology = topology
        self.gpu_aware = gpu_aware      # device buffers go straight into NCCL: always "GPU-aware"

    @property
    def ctx(self):
        return self.child_arch.ctx

    @property
    def backend(self):
        return self.child_arch.backend

    @property
    def device_id(self):
        return self.child_arch.device_id

    @property
    def device(self):
        return self.child_arch.device_id

    def close(self):
        self.child_arch.close()


def Arch(backend, comm=None, dims=None, *, device_id=None, gpu_aware=True) -> Architecture:
    """Arch(backend; device_id=1)                          -- Architectures.jl:46-49
    Arch(backend, comm, dims; device_id, gpu_aware)     -- distributed_architecture.jl:27-34
    (device defaults to the node-local rank + 1)."""
    if comm is None:
        return SingleDeviceArchitecture(backend, 1 if device_id is None else device_id)
    from .distributed import CartesianTopology
    topo = CartesianTopology(comm, tuple(dims))
    dev = topo.shared_rank + 1 if device_id is None else device_id
    child = SingleDeviceArchitecture(backend, dev)
    topo._attach(child)
    return DistributedArchitecture(child, topo, gpu_aware)


def get_backend(arch: Architecture):
    return arch.backend


def get_device(arch: Architecture):
    return arch.device_id


def set_device_(arch_or_dev):
    """set_device!(dev) (ext/ChmyCUDAExt/ChmyCUDAExt.jl:15): every C entry point selects its context's device itself
    (cudaSetDevice at entry), so there is no process-wide current device to switch; returns the device like the
    reference's method does (test/test_architectures.jl:24)."""
    return arch_or_dev.device_id if isinstance(arch_or_dev, Architecture) else arch_or_dev


def is_gpu_aware(arch) -> bool:
    """is_gpu_aware(arch) (distributed_architecture.jl:75): device buffers go straight into NCCL."""
    return bool(getattr(arch, "gpu_aware", True))


def activate_(arch: Architecture, priority: str = "normal"):
    """activate!(arch; priority) (Architectures.jl:71-74).  Stream priorities are fixed inside the context
    (main = normal, boundary = high), so this only validates its argument."""
    if priority not in ("normal", "low", "high"):
        raise ValueError("priority must be :normal, :low or :high")


def synchronize(arch: Architecture):
    """KernelAbstractions.synchronize(backend)."""
    L.check(L.lib().chmy_synchronize(arch.ctx))


def launch_count(arch: Architecture) -> int:
    n = C.c_uint64()
    L.check(L.lib().chmy_ctx_launch_count(arch.ctx, C.byref(n)))
    return int(n.value)


def set_fusion(arch: Architecture, enable=True):
    """`fuse!(arch)`: run `launch(update_stress!)` + `launch(update_velocity!; bc)` (3D) and the 2D / thermal flux -> update
    pairs as ONE sweep each (include/chmy_b200.h: chmy_set_fusion).  Results are identical; the reference has no counterpart.
    True = every fused sweep (3); 1 = only the 3D stress + velocity sweep; False / 0 = off."""
    v = 3 if enable is True else int(enable)
    L.check(L.lib().chmy_set_fusion(arch.ctx, v))


def fused_count(arch: Architecture) -> int:
    n = C.c_uint64(0)
    L.check(L.lib().chmy_fused_count(arch.ctx, C.byref(n)))
    return int(n.value)


def fusion_fallback_count(arch: Architecture) -> int:
    """pairs of launches that ran as two kernels because the device had no memory left for the sweeps' shadow buffers"""
    n = C.c_uint64(0)
    L.check(L.lib().chmy_fusion_fallback_count(arch.ctx, C.byref(n)))
    return int(n.value)


def last_division_mode(arch: Architecture) -> int:
    """how the last fused 3D sweep divided by its uniform scalars: 0 = four operations, 1 = div.rn.f64, 2 = two operations
    (proven exact for all four divisors of the launch, include/chmy_b200.h: chmy_division_two_op_exact)"""
    m = C.c_int32(-1)
    L.check(L.lib().chmy_last_division_mode(arch.ctx, C.byref(m)))
    return int(m.value)


def division_two_op_exact(c: float) -> bool:
    e = C.c_int32(0)
    L.check(L.lib().chmy_division_two_op_exact(float(c), C.byref(e)))
    return bool(e.value)


def overlapped_count(arch: Architecture) -> int:
    """launches whose boundary batches / halo exchange ran behind the boundary tiles of a still-running fused sweep"""
    n = C.c_uint64(0)
    L.check(L.lib().chmy_overlapped_count(arch.ctx, C.byref(n)))
    return int(n.value)


def set_fused_tuning(arch: Architecture, rows_per_cta: int = 0, cluster_size: int = 0, z_chunk: int = 0, variant: int = -1):
    """tile geometry of the fused 3D sweep of THIS context (0 / -1 keep a setting)"""
    L.check(L.lib().chmy_set_fused_tuning(arch.ctx, int(rows_per_cta), int(cluster_size), int(z_chunk), int(variant)))


def set_fused2d_tuning(arch: Architecture, rows_per_chunk: int = 0, unroll: int = 0, thermal3_planes_per_chunk: int = 0):
    L.check(L.lib().chmy_set_fused2d_tuning(arch.ctx, int(rows_per_chunk), int(unroll), int(thermal3_planes_per_chunk)))


def set_launch_split(arch: Architecture, split=True, bc_fold=None):
    """Launches with boundary batches on this context: split truthy ("auto", "on", True; the default) = the batches and the
    halo exchange overlap the kernel (KernelLaunch.jl:160-181: inner region + slabs on two streams; boundary tiles first +
    a retire counter for the fused 3D sweep); False / "off" = one kernel, then the batches, on one stream; "always" = overlap even without a neighbour (tests, A/B).  bc_fold: a batch
    set without exchange as ONE launch (default) or one launch per dimension.  Results are identical
    (include/chmy_b200.h: chmy_set_launch_tuning)."""
    on = 0 if split in (False, 0, "off") else (2 if split in ("always", 2) else 1)
    L.check(L.lib().chmy_set_launch_tuning(arch.ctx, on, -1 if bc_fold is None else int(bool(bc_fold))))


def set_exchange_mode(arch: Architecture, mode="nccl"):
    """Transport of the halo exchange on a distributed architecture: "nccl" (default, the measured path) or "peer"
    (EXPERIMENTAL: the pack kernels store straight into the neighbour's HBM over NVLink, sequence flags instead of
    ncclSend/ncclRecv).  Every rank must choose the same; results are identical (include/chmy_b200.h: chmy_set_exchange_mode)."""
    modes = {"nccl": 0, "peer": 1}
    if mode not in modes:
        raise ValueError(f"exchange mode must be one of {sorted(modes)}, got {mode!r}")
    L.check(L.lib().chmy_set_exchange_mode(arch.ctx, modes[mode]))


def exchange_stats(arch: Architecture):
    """(messages sent as peer stores, messages sent through NCCL) so far on this rank."""
    a, b = C.c_uint64(), C.c_uint64()
    L.check(L.lib().chmy_exchange_stats(arch.ctx, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def topology(arch: DistributedArchitecture):
    return arch.topology


def event_record(arch: Architecture, slot: int):
    """CUDA event on the architecture's main stream (device-side timing for benchmarks)."""
    L.check(L.lib().chmy_event_record(arch.ctx, int(slot)))


def time_fused_sweep(arch: Architecture, slot_begin: int = -1, slot_end: int = -1):
    """record event slot_begin / slot_end right before / after the kernel of every fused 3D sweep (-1, -1: off)"""
    L.check(L.lib().chmy_time_fused_sweep(arch.ctx, int(slot_begin), int(slot_end)))


def event_elapsed_ms(arch: Architecture, start: int, stop: int) -> float:
    ms = C.c_float()
    L.check(L.lib().chmy_event_elapsed_ms(arch.ctx, int(start), int(stop), C.byref(ms)))
    return float(ms.value)
