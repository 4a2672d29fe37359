"""
Launcher: the drop-in boundary.  Mirrors src/KernelLaunch.jl:21-119; `launcher(arch, grid, (op, args); bc=...)`
flattens everything into one POD chmy_launch_desc and makes one call into the C ABI.  The inner/outer split,
streams and events live on the C side (the reference's Workers, src/Workers.jl, are bypassed on this path).
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .boundary_conditions import fill_batch_desc
from .grids import Center, StructuredGrid
from .ops import KernelOp


class Launcher:
    """Launcher(arch, grid; outer_width=nothing) -- KernelLaunch.jl:40-54."""

    def __init__(self, arch, grid: StructuredGrid, *, outer_width=None, blocking: bool = True, exact_split: bool = False):
        self.arch = arch
        self.worksize_ = tuple(n + 2 for n in grid.size(Center()))          # :41
        self.outer_width_ = None if outer_width is None else tuple(int(w) for w in outer_width)
        if self.outer_width_ is not None and len(self.outer_width_) != grid.ndims():
            raise ValueError("outer_width must have one entry per grid dimension")
        # reference semantics: the call returns after the device finished (KernelLaunch.jl:117).  blocking=False
        # keeps everything stream-ordered and defers the wait to `synchronize(arch)` / the next host read.
        self.blocking = blocking
        self.exact_split = exact_split

    # region algebra, KernelLaunch.jl:56-87 (host mirror, used by tests; the C side implements the same formulas)
    def ndims(self):
        return len(self.worksize_)

    def __call__(self, arch, grid: StructuredGrid, kernel_and_args, *, bc=None):
        """launcher(arch, grid, op => args; bc) -- KernelLaunch.jl:105-119."""
        d = self.describe(arch, grid, kernel_and_args, bc=bc)
        L.check(L.lib().chmy_launch(arch.ctx, C.byref(d)))

    def validate(self, grid: StructuredGrid, kernel_and_args, *, bc=None, arch=None):
        """Run chmy_launch's argument checks only (chmy_validate_launch): op id, field count / order / staggered locations
        / sizes / element types, batches.  Needs no device; fields may be `Field.shell`s."""
        d = self.describe(arch, grid, kernel_and_args, bc=bc)
        L.check(L.lib().chmy_validate_launch(C.byref(d)))

    def describe(self, arch, grid: StructuredGrid, kernel_and_args, *, bc=None) -> L.LaunchDesc:
        """Flatten `op => args` (+ bc, outer_width) into the POD chmy_launch_desc."""
        op, args = kernel_and_args
        if not isinstance(op, KernelOp):
            raise TypeError("the B200 path runs the named kernels of chmy_b200.ops, not arbitrary closures")
        fields, scalars, ffield = op.flatten(args)
        d = L.LaunchDesc()
        d.op = op.op_id
        d.oper, d.oper_dim = op.oper, op.oper_dim
        d.flags = (L.LAUNCH_BLOCKING if self.blocking else L.LAUNCH_ASYNC) | (L.LAUNCH_EXACT_SPLIT if self.exact_split else 0)
        d.grid = grid.desc()
        if len(fields) > L.MAX_OP_FIELDS or len(scalars) > L.MAX_SCALARS:
            raise ValueError("too many kernel arguments")
        d.nfields, d.nscalars = len(fields), len(scalars)
        for i, f in enumerate(fields):
            d.fields[i] = None if f is None else f.handle
        for i, s in enumerate(scalars):
            d.scalars[i] = float(s)
        if ffield is not None:
            d.rho_g = ffield.inclusion()
        if bc is not None:
            d.has_bc = 1
            for D, sides in enumerate(bc):
                for S in range(2):
                    fill_batch_desc(d.bc[D][S], sides[S], arch, grid, D, S)
        if self.outer_width_ is not None:
            d.has_outer_width = 1
            for a, w in enumerate(self.outer_width_):
                d.outer_width[a] = w
        return d


def worksize(l: Launcher):
    return l.worksize_


def outer_width(l: Launcher):
    return l.outer_width_


def inner_worksize(l: Launcher):
    return tuple(w - 2 * o for w, o in zip(l.worksize_, l.outer_width_))     # :60


def inner_offset(l: Launcher):
    return l.outer_width_                                                    # :61


def outer_worksize(l: Launcher, D: int):
    """outer_worksize(launcher, Dim(D)), 1-based D -- :63-74."""
    ws, ow = l.worksize_, l.outer_width_
    return tuple(ws[i] if i + 1 < D else ow[i] if i + 1 == D else ws[i] - 2 * ow[i] for i in range(len(ws)))


def outer_offset(l: Launcher, D: int, S: int):
    """outer_offset(launcher, Dim(D), Side(S)), 1-based -- :76-87."""
    ws, ow = l.worksize_, l.outer_width_
    return tuple(0 if i + 1 < D else (0 if S == 1 else ws[i] - ow[i]) if i + 1 == D else ow[i] for i in range(len(ws)))
