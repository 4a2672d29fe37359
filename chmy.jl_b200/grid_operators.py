"""
GridOperators as `launch`-able kernels (include/chmy_b200.h: CHMY_OP_OPERATOR, csrc/operators.cuh).

The reference exposes left/right/δ/∂/∂²/∂k∂, lerp/hlerp, divg/lapl/divg_grad/vmag as point functions that user
`@kernel`s call at an index (src/GridOperators/field_operators.jl:2-125, interpolation.jl:63-94; kernels such as
test/test_grid_operators.jl:24-37).  Arbitrary kernel bodies cannot be JIT-ed through a C ABI, so on this path each
operator is a named kernel `dst[I] = OP(src...)[I]` over the launch range, used like every other op:

    launch(arch, grid, (partial_(1), (dVdx, V.x, grid)))            # ∂x
    launch(arch, grid, (divg_, (C, V, grid)))                         # C[I] = divg(V, grid, I)
    launch(arch, grid, (lerp_, (f_v, f_c, grid)))                     # f_v[I] = lerp(f_c, location(f_v), grid, I)

Dims are 1-based as in the reference (`Dim(1)`).  The library checks that every field sits where the reference's
operator reads / produces it (e.g. ∂ of a Center field along x lives at (Vertex, Center, ...)).
"""
from __future__ import annotations

from . import _lib as L
from .fields import FieldTuple
from .ops import KernelOp


def _t(x):
    return list(x) if isinstance(x, (FieldTuple, tuple, list)) else [x]


def _dst_src(dst, src, g):                       # (dst, f, grid)
    return _t(dst) + _t(src), [], None


def _dst_src_k(dst, src, k, g):                  # (dst, f, k, grid)
    return _t(dst) + _t(src) + [k], [], None


def _dimmed(name, oper, flatten):
    def make(dim: int) -> KernelOp:
        if dim not in (1, 2, 3):
            raise ValueError("dim is 1-based: Dim(1), Dim(2) or Dim(3)")
        return KernelOp(f"{name}(Dim({dim}))", L.OP_OPERATOR, flatten, oper, dim - 1)
    make.__name__ = name
    return make


left_ = _dimmed("left", L.OPER_LEFT, _dst_src)                  # field_operators.jl:2-6
right_ = _dimmed("right", L.OPER_RIGHT, _dst_src)               # field_operators.jl:8-12
delta_ = _dimmed("δ", L.OPER_DELTA, _dst_src)                   # field_operators.jl:14-18
partial_ = _dimmed("∂", L.OPER_PARTIAL, _dst_src)               # field_operators.jl:20-24
partial2_ = _dimmed("∂²", L.OPER_PARTIAL2, _dst_src)            # field_operators.jl:26-30
dkd_ = _dimmed("∂k∂", L.OPER_DKD, _dst_src_k)                   # field_operators.jl:32-36
# cartesian shortcuts, cartesian_field_operators.jl:17-46
dx_, dy_, dz_ = partial_(1), partial_(2), partial_(3)
d2x_, d2y_, d2z_ = partial2_(1), partial2_(2), partial2_(3)

lerp_ = KernelOp("lerp", L.OP_OPERATOR, _dst_src, L.OPER_LERP)                  # interpolation.jl:87
hlerp_ = KernelOp("hlerp", L.OP_OPERATOR, _dst_src, L.OPER_HLERP)               # interpolation.jl:94
divg_ = KernelOp("divg", L.OP_OPERATOR, _dst_src, L.OPER_DIVG)                  # field_operators.jl:50-55
lapl_ = KernelOp("lapl", L.OP_OPERATOR, _dst_src, L.OPER_LAPL)                  # field_operators.jl:72-77
divg_grad_ = KernelOp("divg_grad", L.OP_OPERATOR, _dst_src_k, L.OPER_DIVG_GRAD)  # field_operators.jl:95-100
vmag_ = KernelOp("vmag", L.OP_OPERATOR, _dst_src, L.OPER_VMAG)                  # field_operators.jl:116-121
grad_ = KernelOp("grad", L.OP_OPERATOR, _dst_src, L.OPER_GRAD)                  # test_grid_operators.jl:24-30 (divg1!)
kgrad_ = KernelOp("kgrad", L.OP_OPERATOR, _dst_src_k, L.OPER_KGRAD)             # test_grid_operators.jl:76-82 (divg_grad1!)
