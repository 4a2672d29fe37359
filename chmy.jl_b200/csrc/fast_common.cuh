// fast_common.cuh -- pieces shared by the tuned kernels (ops_fast.cu: 3D, ops_fast2d.cu: 2D).
#pragma once
#include <stdlib.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------- exact division
struct DivC {
    double c, rc, rl;   // divisor, RN(1/c), RN(1/c - rc)
};
// 1 - c * rc is exactly representable when rc is the correctly rounded reciprocal, so the fma returns it without error
static inline DivC divc_of(double c) {
    DivC d;
    d.c = c; d.rc = 1.0 / c; d.rl = fma(-c, d.rc, 1.0) / c;
    return d;
}

static inline bool markstein_ok(double c) {
    unsigned long long b;
    memcpy(&b, &c, sizeof(b));
    const unsigned long long mant = b & 0xFFFFFFFFFFFFFull, ex = (b >> 52) & 0x7FF;
    if (ex == 0 || ex == 0x7FF) return false;            // zero, subnormal, inf, nan
    if (mant == 0xFFFFFFFFFFFFFull) return false;        // the one significand Markstein's theorem excludes
    if (ex < 200 || ex > 1800) return false;             // keep 1/c and the residuals far from under/overflow
    return true;
}

template <bool TRUE_DIV>
__device__ __forceinline__ double div_u(double x, const DivC d) {
    if (TRUE_DIV) return x / d.c;
    // q = RN(x * (rc + rl)) up to 2^-105: within half an ulp (+ that much) of x / c, i.e. a faithful quotient; the residual of a
    // faithful quotient is exact in an fma, and Markstein's theorem (rc = RN(1/c), significand of c not all ones) makes the
    // corrected quotient the correctly rounded one.  Four operations (the reciprocal + two corrections sequence took five).
    const double q = fma(x, d.rc, x * d.rl);
    const double r = fma(-d.c, q, x);
    return fma(r, d.rc, q);
}

// ---------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

#define FULL 0xffffffffu
constexpr int TX = 32;    // lanes per row segment (2 cells each -> 64 cells)
constexpr int TY = 8;     // rows per CTA
constexpr int CZ = 16;    // target z-planes marched by one CTA (measured optimum at 767^3: short chunks, many CTAs)

// element strides of the four (x-location, y-location) storage classes
struct Strides {
    int sy, sz;
};


// fields the tuned kernels can address with aligned 128-bit accesses (they are Float64 kernels: Float32 fields take the
// element-type-generic kernels of ops.cu)
static inline bool aligned16(const chmy_field* f) {
    return f->dtype == CHMY_F64 && f->layout == CHMY_LAYOUT_PITCHED && ((uintptr_t)f->p0 % 16 == 0) && (f->stride[1] % 2 == 0) && (f->stride[2] % 2 == 0) &&
           f->stride[2] * f->sd[2] < (1ll << 40);
}
static inline Strides strides_of(const chmy_field* f) { return Strides{(int)f->stride[1], (int)f->stride[2]}; }
static inline bool same_strides(const chmy_field* a, const chmy_field* b) {
    return a->stride[1] == b->stride[1] && a->stride[2] == b->stride[2];
}

// x-neighbours of a lane's pair of cells (i, i+1): the value at i+2 / i-1 comes from the adjacent lane's register
// unless this lane is the last / first of its warp (or of the box), in which case `edge` holds it (one predicated
// 8-byte load by the caller).  Must be executed by all 32 lanes.
__device__ __forceinline__ double nb_right(double own_x, double edge, bool xlast) {
    const double s = __shfl_down_sync(FULL, own_x, 1);
    return xlast ? edge : s;
}
__device__ __forceinline__ double nb_left(double own_y, double edge, int lane) {
    const double s = __shfl_up_sync(FULL, own_y, 1);
    return lane == 0 ? edge : s;
}

// tuning switches (ops_fast.cu)
bool chmy_fast_disabled();
bool chmy_force_true_div();
// ops_fast2d.cu
int chmy_run_op_fast2d(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st, int* handled);
