"""
The workloads of the hot path, restated against this package's API so that they read like the reference's
example scripts (plots removed):

  diffusion_2d      examples/diffusion_2d.jl:21-54, diffusion_2d_perf.jl:35-69, diffusion_2d_mpi.jl:27-59
  stokes            examples/stokes_2d_inc_ve_T.jl:62-163, stokes_3d_inc_ve_T.jl:79-184,
                    stokes_3d_inc_ve_T_mpi_perf.jl:89-214 (FunctionField rho_g, outer_width, exchange)

Host scalars follow the reference expression by expression.
"""
from __future__ import annotations

import math

import numpy as np

from .architectures import DistributedArchitecture, synchronize
from .boundary_conditions import Dirichlet, Neumann, batch, bc_
from .distributed import allreduce_max
from .fields import Field, FunctionField, TensorField, VectorField, init_incl, interior, maxabs, maxabs_many, set_
from .grids import Center, UniformGrid, Vertex, spacing
from .kernel_launch import Launcher
from .ops import (compute_q_, update_C_, update_old_, update_stress_, update_thermal_, update_thermal_flux_,
                  update_velocity_)



def _sq(x):
    """Julia's x^2 is x*x (Base.literal_pow); Python's x ** 2 is libm pow(x, 2.0), which is not always the same double."""
    return x * x

def _gmax(arch, *vals):
    """max_mpi (stokes_3d_inc_ve_T_mpi_perf.jl:16-19): identity on a single device."""
    if isinstance(arch, DistributedArchitecture):
        return allreduce_max(arch, *vals)
    return vals


# ---------------------------------------------------------------------------------------------- diffusion
class Diffusion2D:
    def __init__(self, arch, nxy, *, outer_width=(16, 8), C0=None, blocking=True, exact_split=False):
        self.arch = arch
        if isinstance(arch, DistributedArchitecture):
            dims_g = tuple(n * p for n, p in zip(nxy, arch.topology.dims))
        else:
            dims_g = tuple(nxy)
        self.grid = grid = UniformGrid(arch, origin=(-1, -1), extent=(2, 2), dims=dims_g)
        self.launch = Launcher(arch, grid, outer_width=outer_width, blocking=blocking, exact_split=exact_split)
        self.chi = 1.0
        self.dt = _sq(min(spacing(grid))) / self.chi / grid.ndims() / 2.1          # diffusion_2d.jl:29
        self.C = Field(arch, grid, Center())
        self.q = VectorField(arch, grid)
        if C0 is not None:
            set_(self.C, C0)                                                       # rand() in the reference (:34)
        bc_(arch, grid, (self.C, Neumann()), exchange=self.C)                      # :35

    def step(self):
        a, g = self.arch, self.grid
        self.launch(a, g, (compute_q_, (self.q, self.C, self.chi, g)))                                      # :46
        self.launch(a, g, (update_C_, (self.C, self.q, self.dt, g)), bc=batch(g, (self.C, Neumann()), exchange=self.C))  # :47

    def run(self, nt):
        for _ in range(nt):
            self.step()
        synchronize(self.arch)

    def fields(self):
        return {"C": self.C, "q.x": self.q.x, "q.y": self.q.y}


# ---------------------------------------------------------------------------------------------- Stokes
class Stokes:
    """Incompressible visco-elastic Stokes + temperature, pseudo-transient (2D or 3D)."""

    def __init__(self, arch, n, *, re_m=2.3 * math.pi, rho_g_function=False, outer_width=None, adv_coef=0.1,
                 blocking=True, exact_split=False):
        self.arch = arch
        N = self.N = len(n)
        self.l = l = (2.0,) * N                                                    # lx, ly, lz
        self.eta, self.G = 1.0e1, 1.0e0
        rho_g = 1.0
        self.psc = self.G
        self.tsc = self.eta / self.psc
        self.T0, self.Ta = 1.0, 0.1
        self.lam = 1e-4 * _sq(l[-1]) / self.tsc                                    # stokes_3d_inc_ve_T.jl:93
        dist = isinstance(arch, DistributedArchitecture)
        dims_g = tuple(a * p for a, p in zip(n, arch.topology.dims)) if dist else tuple(n)
        self.grid = grid = UniformGrid(arch, origin=tuple(-x / 2 for x in l), extent=l, dims=dims_g)
        self.launch = Launcher(arch, grid, outer_width=outer_width, blocking=blocking, exact_split=exact_split)
        self.nx = dims_g[0]
        d = spacing(grid)
        self.d = d
        r = 0.5
        ltau = min(l) / re_m                                                       # :107
        vdt = min(d) / math.sqrt(N * 1.1)                                          # :108
        theta = ltau * (r + 4 / 3) / vdt                                           # :109
        self.dtau_r = 1.0 / (theta + 1.0)                                          # :110
        self.nudtau = vdt * ltau                                                   # :111
        self.dtau_Pr = r / theta                                                   # :112
        self.adv_coef = adv_coef
        A = arch
        self.Pr, self.divV = Field(A, grid, Center()), Field(A, grid, Center())
        self.V, self.r_V = VectorField(A, grid), VectorField(A, grid)
        self.tau, self.tau_old = TensorField(A, grid), TensorField(A, grid)
        self.T, self.T_old = Field(A, grid, Center()), Field(A, grid, Center())
        self.qT = VectorField(A, grid)
        names = ("x0", "y0", "z0")[:N]
        par = {k: 0.0 for k in names}
        rho_loc = tuple(Vertex() if i == N - 1 else Center() for i in range(N))
        if rho_g_function:                                                         # mpi_perf.jl:142 / 2d:107
            self.rho_g = FunctionField(init_incl, grid, rho_loc, parameters={**par, "r": 0.1 * l[0], "in": rho_g, "out": 0.0})
        else:                                                                      # 3d:116,126
            self.rho_g = Field(A, grid, rho_loc)
            set_(self.rho_g, grid, init_incl, parameters={**par, "r": 0.1 * l[0], "in": rho_g, "out": 0.0})
        set_(self.T, grid, init_incl, parameters={**par, "r": 0.1 * l[0], "in": self.T0, "out": self.Ta})
        ax = ("x", "y", "z")[:N]
        Vc = list(self.V)
        self.bc_V = tuple((Vc[i], {a: (Dirichlet() if a == ax[i] else Neumann()) for a in ax}) for i in range(N))
        self.bc_T = ((self.T, Neumann()),)
        self.exch_V = tuple(Vc)
        bc_(A, grid, *self.bc_V, exchange=self.exch_V)                             # :134 / mpi_perf:150
        bc_(A, grid, *self.bc_T, exchange=self.T)                                  # :135
        self.eta_ve = 0.0
        self.dt = 0.0
        self.history = []

    # ---- one outer time step prologue: update_old!, dt, eta_ve  (stokes_3d_inc_ve_T.jl:155-161)
    def begin_time_step(self):
        A, g, N = self.arch, self.grid, self.N
        self.launch(A, g, (update_old_, (self.T, self.tau, self.T_old, self.tau_old)))
        d = self.d
        dt_diff = _sq(min(d)) / self.lam / N / 2.1
        vm = _gmax(A, *maxabs_many(*self.V))                      # one device round trip for the N maxima (:158)
        with np.errstate(divide="ignore"):
            dt_adv = self.adv_coef * min(np.float64(dd) / np.float64(m) for dd, m in zip(d, vm)) / N / 2.1
        self.dt = min(dt_diff, float(dt_adv))
        self.eta_ve = 1.0 / (1.0 / self.eta + 1.0 / (self.G * self.dt))

    def mechanics(self):
        """one PT iteration of the mechanical solver (:164-165)."""
        A, g = self.arch, self.grid
        self.launch(A, g, (update_stress_, (self.tau, self.Pr, self.divV, self.V, self.tau_old, self.eta, self.eta_ve,
                                            self.G, self.dt, self.dtau_Pr, self.dtau_r, g)))
        self.launch(A, g, (update_velocity_, (self.V, self.r_V, self.Pr, self.tau, self.rho_g, self.eta_ve, self.nudtau, g)),
                    bc=batch(g, *self.bc_V, exchange=self.exch_V))

    def thermal(self):
        """thermal sub-step (:167-168)."""
        A, g = self.arch, self.grid
        self.launch(A, g, (update_thermal_flux_, (self.qT, self.T, self.V, self.lam, g)))
        self.launch(A, g, (update_thermal_, (self.T, self.T_old, self.qT, self.dt, g)), bc=batch(g, *self.bc_T, exchange=self.T))

    def residuals(self):
        """:171-175 -- zero the wall-normal residual nodes, then max-norms."""
        A, g, N = self.arch, self.grid, self.N
        ax = ("x", "y", "z")[:N]
        bc_(A, g, *[(rv, {a: Dirichlet()}) for rv, a in zip(self.r_V, ax)])
        loc = maxabs_many(self.divV, *self.r_V)                   # four maxima, one round trip (:172-175)
        glob = _gmax(A, *loc)
        return (glob[0] * self.tsc,) + tuple(x * self.l[-1] / self.psc for x in glob[1:])

    def run(self, nt, niter, ncheck, eps=1e-6, thermal_from_it=2):
        for it in range(1, nt + 1):
            self.begin_time_step()
            for it_pt in range(1, niter + 1):
                self.mechanics()
                if it >= thermal_from_it:
                    self.thermal()
                if it_pt % ncheck == 0:
                    err = self.residuals()
                    self.history.append((it, it_pt) + err)
                    if all(e < eps for e in err):
                        break
                    if not all(math.isfinite(e) for e in err):
                        raise RuntimeError(f"simulation failed, err = {err}")
        synchronize(self.arch)
        return self.history

    def fields(self):
        out = {"Pr": self.Pr, "divV": self.divV, "T": self.T, "T_old": self.T_old}
        for nm, ft in (("V", self.V), ("r_V", self.r_V), ("tau", self.tau), ("tau_old", self.tau_old), ("qT", self.qT)):
            for k in ft.keys():
                out[f"{nm}.{k}"] = getattr(ft, k)
        if isinstance(self.rho_g, Field):
            out["rho_g"] = self.rho_g
        return out
