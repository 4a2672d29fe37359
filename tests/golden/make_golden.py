#!/usr/bin/env python
"""
Generates the committed fixtures under tests/golden/ from the CPU oracle (oracle/).

These are REGRESSION vectors of the oracle -- not outputs of the reference (PTsolvers/Chmy.jl is pure Julia and there
is no Julia toolchain in the build image; the reference's own known-answer tests are re-created directly in
tests/test_oracle_golden.py).  They pin the oracle's bits across hosts/compilers (the -m "not gpu" suite replays the
oracle against them) and give the -m gpu suite a second, stored target for the CUDA path.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    # name: (kind, n, kwargs)
    "stokes3d_10x8x6": ("stokes", (10, 8, 6), dict(rho_g_function=True, nt=2, niter=20, ncheck=5)),
    "stokes3d_field_rho_9x7x5": ("stokes", (9, 7, 5), dict(rho_g_function=False, nt=2, niter=12, ncheck=4)),
    "stokes2d_24x18": ("stokes", (24, 18), dict(rho_g_function=True, nt=2, niter=20, ncheck=5)),
    "diffusion2d_32x24": ("diffusion", (32, 24), dict(nt=10, seed=7)),
    "stokes3d_world2_8x6x6": ("stokes_world", (8, 6, 6), dict(world=2, nt=2, niter=10, ncheck=5)),
}


def run_case(kind, n, kw):
    import drivers as OD
    import oracle as o
    out = {}
    if kind == "stokes":
        s = OD.Stokes(n, rho_g_function=kw["rho_g_function"])
        h = s.run(kw["nt"], kw["niter"], kw["ncheck"])
        out["history"] = np.array(h, dtype=np.float64)
        out["dt_eta_ve"] = np.array([s.dt, s.eta_ve])
        for k, f in s.fields().items():
            out["f:" + k] = f.data.copy()
    elif kind == "diffusion":
        C0 = np.random.default_rng(kw["seed"]).random(n)
        s = OD.Diffusion2D(n, C0=C0)
        s.run(kw["nt"])
        for k, f in s.fields().items():
            out["f:" + k] = f.data.copy()
    elif kind == "stokes_world":
        pd = o.dims_create(kw["world"], (0,) * len(n))
        s = OD.Stokes(n, proc_dims=pd, rho_g_function=True, outer_width=(3,) * len(n), adv_coef=0.01, re_m=2.5 * np.pi)
        h = s.run(kw["nt"], kw["niter"], kw["ncheck"])
        out["history"] = np.array(h, dtype=np.float64)
        for r in range(kw["world"]):
            for k, f in s.fields(r).items():
                out[f"r{r}:{k}"] = f.data.copy()
    return out


def halo_case():
    """pack buffers and unpacked halos of index-encoded fields (communication_views.jl:1-34)."""
    import oracle as o
    out = {}
    for n, loc in (((9, 6), (1, 0)), ((7, 5, 4), (0, 1, 1)), ((7, 5, 4), (1, 0, 0))):
        g = o.Grid((-1.0,) * len(n), (2.0,) * len(n), n)
        f = o.Field(g, loc)
        f.data[...] = np.arange(f.data.size, dtype=np.float64).reshape(f.sdims, order="F")
        tag = "x".join(map(str, n)) + "_" + "".join(map(str, loc))
        for D in range(len(n)):
            for S in range(2):
                out[f"pack:{tag}:{D}{S}"] = o.pack_send(f, D, S).copy()
    return out


def main():
    for name, (kind, n, kw) in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(kind, n, kw))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "halo_pack.npz"), **halo_case())
    print("wrote halo_pack")


if __name__ == "__main__":
    main()
