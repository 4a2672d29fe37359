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

// Is  fma(x, rc, RN(x * rl))  -- the double-double reciprocal product alone, TWO operations -- the correctly rounded x / c for
// EVERY normal x?  Decidable per divisor (after Brisebarre & Muller, "Correctly rounded multiplication by arbitrary precision
// constants"): write c = M * 2^a with M odd and x = X * 2^b with a 53-bit integer X; scaled so that one ulp of the quotient is 1
// the quotient is Y = X * 2^s / M (s = bits(M) or bits(M) - 1), and the value the fma rounds differs from Y by less than
// 2^-52 (1 + 2^-51) (the errors of rl and of RN(x * rl)).  The result can only be wrong if a rounding midpoint n + 1/2 lies that
// close to Y, i.e. if |2^(s+1) X - (2n+1) M| <= 2 M 2^-51 = M 2^-50 < 8: at most eight odd residues N, each a congruence
// 2^(s+1) X = N (mod M) with a handful of solutions X in [2^52, 2^53) when M has 51+ bits and none needed otherwise.  Every
// solution is tried against IEEE division here, on the host (fma is exactly specified, the device computes the same bits):
// if none fails, no operand can.  About 1.3 % of random divisors do have a failing operand (which random testing never finds);
// they keep the four-operation sequence below.
static inline bool div2_exact(double c) {
    if (!markstein_ok(c)) return false;
    c = fabs(c);
    int e;
    const double f = frexp(c, &e);
    unsigned long long M = (unsigned long long)ldexp(f, 53);
    while ((M & 1ull) == 0) M >>= 1;
    if (M == 1) return true;                                   // a power of two: every step is exact
    const unsigned long long K = M >> 50;                      // |N| <= K
    if (K == 0) return true;
    const int L = 64 - __builtin_clzll(M);
    const DivC d = divc_of(c);
    const unsigned long long inv2 = (M + 1) >> 1, lo = 1ull << 52, hi = 1ull << 53;
    for (int s = L - 1; s <= L; ++s) {
        unsigned long long inv = 1;                            // 2^-(s+1) mod M
        for (int t = 0; t <= s; ++t) inv = (unsigned long long)((unsigned __int128)inv * inv2 % M);
        for (unsigned long long N = 1; N <= K; N += 2)
            for (int sg = 0; sg < 2; ++sg) {
                const unsigned long long r = sg ? M - N % M : N % M;
                const unsigned long long X0 = (unsigned long long)((unsigned __int128)(r % M) * inv % M);
                for (unsigned long long X = X0 >= lo ? X0 : X0 + (lo - X0 + M - 1) / M * M; X < hi; X += M) {
                    const double x = (double)X;
                    if (fma(x, d.rc, x * d.rl) != x / c) return false;
                }
            }
    }
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

// division mode of the fused 3D sweep: 0 = the four-operation sequence, 1 = div.rn.f64, 2 = the two-operation sequence (only when
// div2_exact() holds for every divisor of the launch)
template <int MODE>
__device__ __forceinline__ double div_m(double x, const DivC d) {
    if (MODE == 2) return fma(x, d.rc, x * d.rl);
    return div_u<MODE == 1>(x, d);
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
