"""
The kernels `launch` can run on this path: the `@kernel` functions of the named example solvers, identified by
function object (as the Julia glue would dispatch on `typeof(op)`), each with the flattening of its argument tuple
into (device fields, scalars, optional FunctionField) in the order include/chmy_b200.h documents.
Argument tuples are exactly the reference's, including the trailing `grid`:
    examples/diffusion_2d.jl:46-47, examples/stokes_3d_inc_ve_T.jl:155-168.
"""
from __future__ import annotations

from . import _lib as L
from .fields import ConstantField, Field, FieldTuple, FunctionField


class KernelOp:
    def __init__(self, name: str, op_id: int, flatten, oper: int = 0, oper_dim: int = 0):
        self.name, self.op_id, self._flatten = name, op_id, flatten
        self.oper, self.oper_dim = oper, oper_dim       # CHMY_OP_OPERATOR only (grid_operators.py)

    def flatten(self, args):
        return self._flatten(*args)

    def __repr__(self):
        return self.name


def _t(x):
    return list(x) if isinstance(x, FieldTuple) else [x]


def _compute_q(q, C, chi, g):
    return _t(q) + [C], [chi], None


def _update_C(C, q, dt, g):
    return [C] + _t(q), [dt], None


def _update_old(T, tau, T_old, tau_old):
    return [T] + _t(tau) + [T_old] + _t(tau_old), [], None


def _update_stress(tau, Pr, divV, V, tau_old, eta, eta_ve, G, dt, dtau_Pr, dtau_r, g):
    return _t(tau) + [Pr, divV] + _t(V) + _t(tau_old), [eta, eta_ve, G, dt, dtau_Pr, dtau_r], None


class _InclusionOf:
    """adapter: anything that can hand the launch a chmy_inclusion"""

    def __init__(self, inc):
        self._inc = inc

    def inclusion(self):
        return self._inc


def _update_velocity(V, r_V, Pr, tau, rhog, eta_ve, nudtau, g):
    if isinstance(rhog, ConstantField):           # ZeroField / OneField / ValueField: a constant body force
        N = g.ndims()
        loc = _t(V)[N - 1].loc                    # rho_g lives where the last velocity component lives
        return _t(V) + _t(r_V) + [Pr] + _t(tau) + [None], [eta_ve, nudtau], _InclusionOf(rhog.inclusion_at(g, loc))
    if isinstance(rhog, FunctionField) and not rhog.in_kernel():
        rhog = rhog.materialize(Pr.arch)          # any other function body: host-evaluated once into a stored Field
    if isinstance(rhog, FunctionField):
        return _t(V) + _t(r_V) + [Pr] + _t(tau) + [None], [eta_ve, nudtau], rhog
    return _t(V) + _t(r_V) + [Pr] + _t(tau) + [rhog], [eta_ve, nudtau], None


def _update_thermal_flux(qT, T, V, lam, g):
    return _t(qT) + [T] + _t(V), [lam], None


def _update_thermal(T, T_old, qT, dt, g):
    return [T, T_old] + _t(qT), [dt], None


compute_q_ = KernelOp("compute_q!", L.OP_COMPUTE_Q, _compute_q)
update_C_ = KernelOp("update_C!", L.OP_UPDATE_C, _update_C)
update_old_ = KernelOp("update_old!", L.OP_UPDATE_OLD, _update_old)
update_stress_ = KernelOp("update_stress!", L.OP_UPDATE_STRESS, _update_stress)
update_velocity_ = KernelOp("update_velocity!", L.OP_UPDATE_VELOCITY, _update_velocity)
update_thermal_flux_ = KernelOp("update_thermal_flux!", L.OP_UPDATE_THERMAL_FLUX, _update_thermal_flux)
update_thermal_ = KernelOp("update_thermal!", L.OP_UPDATE_THERMAL, _update_thermal)
