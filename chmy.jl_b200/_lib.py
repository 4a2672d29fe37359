"""
ctypes binding of libchmy_b200.so (include/chmy_b200.h).  This is the binding a Julia `ChmyB200Ext` would write
with `ccall`; here it is the Python stand-in for the absent Julia toolchain.

There is deliberately NO fallback: if the CUDA library is missing or no B200 is visible, every compute entry point
raises.  Nothing in this package imports or calls the CPU oracle under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libchmy_b200.so")

MAX_DIMS, MAX_BATCH_FIELDS, MAX_OP_FIELDS, MAX_SCALARS, UNIQUE_ID_BYTES = 3, 8, 24, 8, 128

CENTER, VERTEX = 0, 1
BOUNDED, CONNECTED = 0, 1
DIRICHLET, NEUMANN = 0, 1
BATCH_EMPTY, BATCH_FIELD, BATCH_EXCHANGE = 0, 1, 2
LAYOUT_PITCHED, LAYOUT_DENSE = 0, 1
F64, F32 = 0, 1
LAUNCH_ASYNC, LAUNCH_BLOCKING, LAUNCH_EXACT_SPLIT = 0, 1, 2

OP_NONE, OP_COMPUTE_Q, OP_UPDATE_C, OP_UPDATE_OLD, OP_UPDATE_STRESS, OP_UPDATE_VELOCITY, OP_UPDATE_THERMAL_FLUX, \
    OP_UPDATE_THERMAL, OP_OPERATOR = range(9)
# chmy_operator
OPER_LEFT, OPER_RIGHT, OPER_DELTA, OPER_PARTIAL, OPER_PARTIAL2, OPER_DKD, OPER_LERP, OPER_HLERP, OPER_DIVG, OPER_LAPL, \
    OPER_DIVG_GRAD, OPER_VMAG, OPER_GRAD, OPER_KGRAD = range(1, 15)


class ChmyError(RuntimeError):
    """Non-zero status from the C ABI (the Julia glue would `error(msg)`)."""


class GridDesc(C.Structure):
    _fields_ = [("ndims", C.c_int32), ("_pad", C.c_int32), ("n", C.c_int64 * MAX_DIMS),
                ("origin", C.c_double * MAX_DIMS), ("extent", C.c_double * MAX_DIMS),
                ("spacing", C.c_double * MAX_DIMS), ("inv_spacing", C.c_double * MAX_DIMS),
                ("connectivity", (C.c_int32 * 2) * MAX_DIMS)]


class BatchDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nfields", C.c_int32), ("fields", C.c_void_p * MAX_BATCH_FIELDS),
                ("bc_kind", C.c_int32 * MAX_BATCH_FIELDS), ("value", C.c_double * MAX_BATCH_FIELDS),
                ("value_field", C.c_void_p * MAX_BATCH_FIELDS)]


class Inclusion(C.Structure):
    _fields_ = [("active", C.c_int32), ("loc", C.c_int32 * MAX_DIMS), ("c0", C.c_double * MAX_DIMS),
                ("r", C.c_double), ("inn", C.c_double), ("out", C.c_double)]


class LaunchDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("flags", C.c_int32), ("grid", GridDesc),
                ("nfields", C.c_int32), ("nscalars", C.c_int32),
                ("fields", C.c_void_p * MAX_OP_FIELDS), ("scalars", C.c_double * MAX_SCALARS),
                ("rho_g", Inclusion), ("has_bc", C.c_int32), ("has_outer_width", C.c_int32),
                ("outer_width", C.c_int64 * MAX_DIMS), ("bc", (BatchDesc * 2) * MAX_DIMS),
                ("oper", C.c_int32), ("oper_dim", C.c_int32)]


class FieldInfo(C.Structure):
    _fields_ = [("ndims", C.c_int32), ("layout", C.c_int32), ("loc", C.c_int32 * MAX_DIMS),
                ("dims", C.c_int64 * MAX_DIMS), ("stride", C.c_int64 * MAX_DIMS),
                ("origin_ptr", C.c_void_p), ("base_ptr", C.c_void_p), ("bytes", C.c_size_t),
                ("dtype", C.c_int32), ("_pad", C.c_int32)]


# every symbol include/chmy_b200.h declares: (name, restype, argtypes)
_P = C.POINTER
_i64p, _i32p, _dp, _vp = _P(C.c_int64), _P(C.c_int32), _P(C.c_double), C.c_void_p
SYMBOLS = {
    "chmy_abi_version": (C.c_int, []),
    "chmy_last_error": (C.c_char_p, []),
    "chmy_device_count": (C.c_int, [_P(C.c_int)]),
    "chmy_struct_size": (C.c_size_t, [C.c_int]),
    "chmy_ctx_create": (C.c_int, [C.c_int, _P(_vp)]),
    "chmy_ctx_destroy": (C.c_int, [_vp]),
    "chmy_ctx_device": (C.c_int, [_vp, _P(C.c_int)]),
    "chmy_synchronize": (C.c_int, [_vp]),
    "chmy_ctx_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "chmy_event_record": (C.c_int, [_vp, C.c_int]),
    "chmy_event_elapsed_ms": (C.c_int, [_vp, C.c_int, C.c_int, _P(C.c_float)]),
    "chmy_time_fused_sweep": (C.c_int, [_vp, C.c_int, C.c_int]),
    "chmy_ctx_streams": (C.c_int, [_vp, _P(_vp), _P(_vp)]),
    "chmy_dims_create": (C.c_int, [C.c_int, C.c_int, _i32p]),
    "chmy_comm_unique_id": (C.c_int, [_P(C.c_uint8)]),
    "chmy_topo_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _i32p, _P(C.c_uint8)]),
    "chmy_topo_coords": (C.c_int, [_vp, _i32p]),
    "chmy_topo_neighbors": (C.c_int, [_vp, _P(C.c_int32 * 2)]),
    "chmy_allreduce_max": (C.c_int, [_vp, _dp, C.c_int]),
    "chmy_barrier": (C.c_int, [_vp]),
    "chmy_field_create": (C.c_int, [_vp, C.c_int, _i64p, _i32p, C.c_int, _P(_vp)]),
    "chmy_field_create_typed": (C.c_int, [_vp, C.c_int, _i64p, _i32p, C.c_int, C.c_int, _P(_vp)]),
    "chmy_field_create_shell": (C.c_int, [C.c_int, _i64p, _i32p, C.c_int, C.c_int, _P(_vp)]),
    "chmy_field_destroy": (C.c_int, [_vp]),
    "chmy_field_get_info": (C.c_int, [_vp, _P(FieldInfo)]),
    "chmy_field_fill": (C.c_int, [_vp, _vp, C.c_double, _i64p, _i64p]),
    "chmy_field_copy_from_host": (C.c_int, [_vp, _vp, _vp, _i64p, _i64p]),
    "chmy_field_copy_to_host": (C.c_int, [_vp, _vp, _vp, _i64p, _i64p]),
    "chmy_field_copy": (C.c_int, [_vp, _vp, _vp, _i64p, _i64p]),
    "chmy_field_set_inclusion": (C.c_int, [_vp, _vp, _P(GridDesc), _P(Inclusion)]),
    "chmy_field_set_gaussian": (C.c_int, [_vp, _vp, _P(GridDesc)]),
    "chmy_field_maxabs": (C.c_int, [_vp, _vp, _i64p, _i64p, _dp]),
    "chmy_field_maxabs_many": (C.c_int, [_vp, C.c_int, _P(_vp), _i64p, _i64p, _dp]),
    "chmy_host_alloc": (C.c_int, [_vp, C.c_size_t, _P(_vp)]),
    "chmy_host_free": (C.c_int, [_vp, _vp]),
    "chmy_launch": (C.c_int, [_vp, _P(LaunchDesc)]),
    "chmy_validate_launch": (C.c_int, [_P(LaunchDesc)]),
    "chmy_bc": (C.c_int, [_vp, _P(GridDesc), _P(BatchDesc * 2), C.c_int]),
    "chmy_exchange_halo": (C.c_int, [_vp, _P(GridDesc), C.c_int, C.c_int, C.c_int, _P(_vp), C.c_int]),
    "chmy_exchange_halo_all": (C.c_int, [_vp, _P(GridDesc), C.c_int, _P(_vp), C.c_int]),
    "chmy_set_exchange_mode": (C.c_int, [_vp, C.c_int]),
    "chmy_exchange_stats": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_uint64)]),
    "chmy_selftest_division": (C.c_int, [_vp, C.c_double, C.c_longlong, C.c_ulonglong, _P(C.c_ulonglong), _P(C.c_int)]),
    "chmy_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "chmy_set_launch_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "chmy_overlapped_count": (C.c_int, [C.c_void_p, _P(C.c_uint64)]),
    "chmy_launch_split_plan": (C.c_int, [_P(LaunchDesc), _i32p, C.c_int32, _P(C.c_int32), _i32p, _i32p]),
    "chmy_selftest_tile_order": (C.c_int, [_i32p, _i32p, _i32p, C.c_int32, _i32p]),
    "chmy_set_fusion": (C.c_int, [_vp, C.c_int]),
    "chmy_fused_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "chmy_fusion_fallback_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "chmy_division_two_op_exact": (C.c_int, [C.c_double, _P(C.c_int32)]),
    "chmy_last_division_mode": (C.c_int, [_vp, _P(C.c_int32)]),
    "chmy_selftest_division2": (C.c_int, [_vp, C.c_double, C.c_longlong, C.c_ulonglong, _P(C.c_ulonglong), _P(C.c_int)]),
    "chmy_set_fused_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "chmy_set_fused2d_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "chmy_halo_slab_len": (C.c_int, [_vp, C.c_int, _i64p]),
    "chmy_halo_pack": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "chmy_halo_unpack": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
}

ABI_VERSION = 3
_lib = None


def lib():
    """Load libchmy_b200.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ChmyError(f"{LIB_PATH} is missing: build it with `python chmy.jl_b200/build.py` "
                            "(there is no CPU fallback on this path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError here == ABI mismatch: fail loudly
            fn.restype, fn.argtypes = res, args
        if L.chmy_abi_version() != ABI_VERSION:
            raise ChmyError("libchmy_b200.so ABI version mismatch")
        for which, st in enumerate((GridDesc, BatchDesc, Inclusion, LaunchDesc, FieldInfo)):
            if L.chmy_struct_size(which) != C.sizeof(st):
                raise ChmyError(f"struct layout mismatch for {st.__name__}: library {L.chmy_struct_size(which)}, "
                                f"binding {C.sizeof(st)}")
        _lib = L
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().chmy_last_error().decode("utf-8", "replace")
        raise ChmyError(f"chmy_b200 error {status}: {msg}")


def i64x3(vals, fill=0):
    vals = [int(v) for v in vals]
    return (C.c_int64 * 3)(*(vals + [fill] * (3 - len(vals))))


def i32x3(vals, fill=0):
    vals = [int(v) for v in vals]
    return (C.c_int32 * 3)(*(vals + [fill] * (3 - len(vals))))
