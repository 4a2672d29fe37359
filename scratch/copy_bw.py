"""Host<->device copy bandwidth on this box: contiguous vs what set!/interior do (scratch measurement)."""
import time, torch, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = 767
nbytes = n * n * n * 8
h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
for name, fn in (("H2D contiguous", lambda: d.copy_(h, non_blocking=True)), ("D2H contiguous", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2):
        fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {nbytes / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms for {nbytes / 1e9:.2f} GB)", flush=True)
# chunked contiguous copies (64 MB pieces), as a staged path would issue them
ch = 64 << 20
def chunked(dst, src):
    for o in range(0, nbytes // 8, ch // 8):
        dst[o:o + ch // 8].copy_(src[o:o + ch // 8], non_blocking=True)
for name, fn in (("H2D 64MB chunks", lambda: chunked(d, h)), ("D2H 64MB chunks", lambda: chunked(h, d))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {nbytes / dt / 1e9:.1f} GB/s", flush=True)
import chmy_b200 as chy
arch = chy.Arch(chy.B200Backend())
g = chy.UniformGrid(arch, origin=(0, 0, 0), extent=(1, 1, 1), dims=(n, n, n))
f = chy.Field(arch, g, chy.Center())
v = chy.pinned_array(arch, f.dims)
v[...] = 1.0
for name, fn in (("set!(f, A_host)", lambda: chy.set_(f, v)), ("Array(interior(f))", lambda: chy.interior(f, out=v))):
    fn(); chy.synchronize(arch)
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    chy.synchronize(arch)
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {nbytes / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)", flush=True)
arch.close()
