"""
Oracle-side drivers of the five configurations (TEST INFRASTRUCTURE ONLY): the same workloads as
chmy.jl_b200/drivers.py, written against oracle.py and able to simulate a whole Cartesian process grid in one
process (ranks advanced in lock step).  References: examples/diffusion_2d*.jl, examples/stokes_{2,3}d_inc_ve_T*.jl.
"""
from __future__ import annotations

import math

import numpy as np

import oracle as o



def _sq(x):
    """Julia's x^2 is x*x (Base.literal_pow); Python's x ** 2 is libm pow(x, 2.0), which is not always the same double."""
    return x * x

def _world(n_local, proc_dims):
    nd = len(n_local)
    P = int(np.prod(proc_dims)) if proc_dims else 1
    proc_dims = tuple(proc_dims) if proc_dims else (1,) * nd
    topos = [o.Topology(P, proc_dims, r) for r in range(P)]
    return P, proc_dims, topos


class Diffusion2D:
    def __init__(self, nxy, proc_dims=None, outer_width=(16, 8), C0=None):
        self.P, pd, self.topos = _world(nxy, proc_dims)
        dims_g = tuple(n * p for n, p in zip(nxy, pd))
        self.grids = [o.local_grid((-1.0, -1.0), (2.0, 2.0), dims_g, t) for t in self.topos]
        self.launchers = [o.Launcher(g, outer_width) for g in self.grids]
        self.chi = 1.0
        g0 = self.grids[0]
        self.dt = _sq(min(g0.spacing)) / self.chi / 2 / 2.1
        self.C = [o.Field(g, o.CENTER) for g in self.grids]
        self.q = [o.VectorField(g) for g in self.grids]
        if C0 is not None:
            for r in range(self.P):
                self.C[r].set(C0[r] if isinstance(C0, (list, tuple)) else C0)
        self._bcs = lambda: [o.batch(self.grids[r], (self.C[r], o.Neumann()), exchange=self.C[r]) for r in range(self.P)]
        o.bc_world(self.grids, self._bcs(), self.topos)

    def step(self):
        o.launch_world(self.launchers, self.grids, o.compute_q, [(self.q[r], self.C[r], self.chi) for r in range(self.P)])
        o.launch_world(self.launchers, self.grids, o.update_C, [(self.C[r], self.q[r], self.dt) for r in range(self.P)],
                       self._bcs(), self.topos)

    def run(self, nt):
        for _ in range(nt):
            self.step()

    def fields(self, r=0):
        return {"C": self.C[r], "q.x": self.q[r]["x"], "q.y": self.q[r]["y"]}


class Stokes:
    def __init__(self, n, proc_dims=None, re_m=2.3 * math.pi, rho_g_function=False, outer_width=None, adv_coef=0.1):
        N = self.N = len(n)
        self.P, pd, self.topos = _world(n, proc_dims)
        P = self.P
        self.l = l = (2.0,) * N
        self.eta, self.G = 1.0e1, 1.0e0
        rho_g = 1.0
        self.psc = self.G
        self.tsc = self.eta / self.psc
        self.T0, self.Ta = 1.0, 0.1
        self.lam = 1e-4 * _sq(l[-1]) / self.tsc
        dims_g = tuple(a * p for a, p in zip(n, pd))
        self.grids = [o.local_grid(tuple(-x / 2 for x in l), l, dims_g, t) for t in self.topos]
        self.launchers = [o.Launcher(g, outer_width) for g in self.grids]
        self.nx = dims_g[0]
        d = self.d = self.grids[0].spacing
        r = 0.5
        ltau = min(l) / re_m
        vdt = min(d) / math.sqrt(N * 1.1)
        theta = ltau * (r + 4 / 3) / vdt
        self.dtau_r = 1.0 / (theta + 1.0)
        self.nudtau = vdt * ltau
        self.dtau_Pr = r / theta
        self.adv_coef = adv_coef
        G = self.grids
        self.Pr = [o.Field(g, o.CENTER) for g in G]
        self.divV = [o.Field(g, o.CENTER) for g in G]
        self.V = [o.VectorField(g) for g in G]
        self.r_V = [o.VectorField(g) for g in G]
        self.tau = [o.TensorField(g) for g in G]
        self.tau_old = [o.TensorField(g) for g in G]
        self.T = [o.Field(g, o.CENTER) for g in G]
        self.T_old = [o.Field(g, o.CENTER) for g in G]
        self.qT = [o.VectorField(g) for g in G]
        rho_loc = tuple(o.VERTEX if i == N - 1 else o.CENTER for i in range(N))
        c0 = (0.0,) * N
        if rho_g_function:
            self.rho_g = [o.Inclusion(rho_loc, c0, 0.1 * l[0], rho_g, 0.0) for _ in G]
        else:
            self.rho_g = [o.Field(g, rho_loc) for g in G]
            for f in self.rho_g:
                o.set_inclusion(f, o.Inclusion(rho_loc, c0, 0.1 * l[0], rho_g, 0.0))
        for f in self.T:
            o.set_inclusion(f, o.Inclusion(f.loc, c0, 0.1 * l[0], self.T0, self.Ta))
        self.ax = o.AXES[:N]
        o.bc_world(G, self._bc_V(), self.topos)
        o.bc_world(G, self._bc_T(), self.topos)
        self.eta_ve = 0.0
        self.dt = 0.0
        self.history = []

    def _bc_V(self):
        out = []
        for r in range(self.P):
            V = self.V[r]
            specs = [(V[a], {b: (o.Dirichlet() if a == b else o.Neumann()) for b in self.ax}) for a in self.ax]
            out.append(o.batch(self.grids[r], *specs, exchange=tuple(V[a] for a in self.ax)))
        return out

    def _bc_T(self):
        return [o.batch(self.grids[r], (self.T[r], o.Neumann()), exchange=self.T[r]) for r in range(self.P)]

    def begin_time_step(self):
        P, N = self.P, self.N
        o.launch_world(self.launchers, self.grids, o.update_old,
                       [(self.T[r], self.tau[r], self.T_old[r], self.tau_old[r]) for r in range(P)])
        d = self.d
        dt_diff = _sq(min(d)) / self.lam / N / 2.1
        vm = [max(self.V[r][a].maxabs() for r in range(P)) for a in self.ax]
        with np.errstate(divide="ignore"):
            dt_adv = self.adv_coef * min(np.float64(dd) / np.float64(m) for dd, m in zip(d, vm)) / N / 2.1
        self.dt = min(dt_diff, float(dt_adv))
        self.eta_ve = 1.0 / (1.0 / self.eta + 1.0 / (self.G * self.dt))

    def mechanics(self):
        P = self.P
        o.launch_world(self.launchers, self.grids, o.update_stress,
                       [(self.tau[r], self.Pr[r], self.divV[r], self.V[r], self.tau_old[r], self.eta, self.eta_ve, self.G,
                         self.dt, self.dtau_Pr, self.dtau_r) for r in range(P)])
        o.launch_world(self.launchers, self.grids, o.update_velocity,
                       [(self.V[r], self.r_V[r], self.Pr[r], self.tau[r], self.rho_g[r], self.eta_ve, self.nudtau)
                        for r in range(P)], self._bc_V(), self.topos)

    def thermal(self):
        P = self.P
        o.launch_world(self.launchers, self.grids, o.update_thermal_flux,
                       [(self.qT[r], self.T[r], self.V[r], self.lam) for r in range(P)])
        o.launch_world(self.launchers, self.grids, o.update_thermal,
                       [(self.T[r], self.T_old[r], self.qT[r], self.dt) for r in range(P)], self._bc_T(), self.topos)

    def residuals(self):
        P = self.P
        bcs = [o.batch(self.grids[r], *[(self.r_V[r][a], {a: o.Dirichlet()}) for a in self.ax]) for r in range(P)]
        o.bc_world(self.grids, bcs, self.topos)
        e0 = max(self.divV[r].maxabs() for r in range(P)) * self.tsc
        rest = tuple(max(self.r_V[r][a].maxabs() for r in range(P)) * self.l[-1] / self.psc for a in self.ax)
        return (e0,) + rest

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
        return self.history

    def fields(self, r=0):
        out = {"Pr": self.Pr[r], "divV": self.divV[r], "T": self.T[r], "T_old": self.T_old[r]}
        for nm, ft in (("V", self.V[r]), ("r_V", self.r_V[r]), ("tau", self.tau[r]), ("tau_old", self.tau_old[r]),
                       ("qT", self.qT[r])):
            for k, f in ft.items():
                out[f"{nm}.{k}"] = f
        if isinstance(self.rho_g[r], o.Field):
            out["rho_g"] = self.rho_g[r]
        return out
