// operators.cuh -- the staggered-grid operators of src/GridOperators as field-level kernels.
//
// The reference exposes left/right/δ/∂/∂²/∂k∂, itp/lerp/hlerp, divg/lapl/divg_grad/vmag as point functions that user
// `@kernel`s call at an index I (test/test_grid_operators.jl:21-126, test/test_interpolations.jl:22-70).  A C library
// cannot JIT such kernels, so this path offers the operators one level up: `dst[I] = OP(src...)[I]` over the launch
// range, selected by chmy_launch_desc::oper.  The point functions below are plain C++ shared by nvcc (device code) and
// the host compiler (tests/emul/operators_emul.cpp), which proves them bit-identical to the oracle without a GPU.
//
// Arithmetic contract as everywhere on this path: every + - * / is one IEEE operation of the field's element type T
// (Float64 | Float32, test/common.jl:9) in the reference's evaluation order (compiled with -fmad=false /
// -ffp-contract=off); fma only where the reference writes muladd (interpolation.jl:14,15).  Uniform grids:
// iΔ(grid, loc, dim, I) is the stored inv_spacing, itp weights are convert(T, 0.5) (interpolation.jl:29-33).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define OPR_HD __host__ __device__ __forceinline__
#else
#define OPR_HD inline
#endif

#define OPR_CENTER 0
#define OPR_VERTEX 1

// operator ids == chmy_operator (include/chmy_b200.h)
enum {
    OPR_LEFT = 1, OPR_RIGHT = 2, OPR_DELTA = 3, OPR_PARTIAL = 4, OPR_PARTIAL2 = 5, OPR_DKD = 6, OPR_LERP = 7,
    OPR_HLERP = 8, OPR_DIVG = 9, OPR_LAPL = 10, OPR_DIVG_GRAD = 11, OPR_VMAG = 12, OPR_GRAD = 13, OPR_KGRAD = 14
};

// A field as the operators see it: p addresses logical index 0 of every active dim, x stride 1.
template <class T>
struct OprFieldT {
    T*        p;
    long long sy, sz;
    int       loc[3];
};
using OprField = OprFieldT<double>;

template <class T>
OPR_HD T opr_at(const OprFieldT<T>& f, const int* I) {
    return f.p[(long long)I[0] + (long long)I[1] * f.sy + (long long)I[2] * f.sz];
}

// GridOperators.jl:23-36 with field_operators.jl:2-12 (from = flipped(loc, dim)):
//   f Vertex along dim: left = f[I],   right = f[I+e];     f Center along dim: left = f[I-e], right = f[I]
template <class T>
OPR_HD T opr_left(const OprFieldT<T>& f, int dim, const int* I) {
    int J[3] = {I[0], I[1], I[2]};
    if (f.loc[dim] == OPR_CENTER) J[dim] -= 1;
    return opr_at(f, J);
}
template <class T>
OPR_HD T opr_right(const OprFieldT<T>& f, int dim, const int* I) {
    int J[3] = {I[0], I[1], I[2]};
    if (f.loc[dim] == OPR_VERTEX) J[dim] += 1;
    return opr_at(f, J);
}
// partial_derivatives.jl:2,5
template <class T>
OPR_HD T opr_delta(const OprFieldT<T>& f, int dim, const int* I) { return opr_right(f, dim, I) - opr_left(f, dim, I); }
template <class T>
OPR_HD T opr_partial(const OprFieldT<T>& f, const T* id, int dim, const int* I) { return opr_delta(f, dim, I) * id[dim]; }

// partial_derivatives.jl:7-12 with field_operators.jl:26-30 (from = loc): Ir = ir(flip(L), L), Il = il(flip(L), L)
//   f Center along dim: Ir = I+e, Il = I;      f Vertex along dim: Ir = I, Il = I-e
template <class T>
OPR_HD void opr_second_idx(const OprFieldT<T>& f, int dim, const int* I, int* Ir, int* Il) {
    for (int a = 0; a < 3; ++a) Ir[a] = Il[a] = I[a];
    if (f.loc[dim] == OPR_CENTER) Ir[dim] += 1; else Il[dim] -= 1;
}
template <class T>
OPR_HD T opr_partial2(const OprFieldT<T>& f, const T* id, int dim, const int* I) {
    int Ir[3], Il[3];
    opr_second_idx(f, dim, I, Ir, Il);
    return (opr_partial(f, id, dim, Ir) - opr_partial(f, id, dim, Il)) * id[dim];
}

// interpolation.jl:14-15
template <class T>
OPR_HD T opr_rule(bool harmonic, T t, T a, T b) {
    if (!harmonic) return fma(t, b - a, a);
    const T ia = (T)1.0 / a, ib = (T)1.0 / b;
    return (T)1.0 / fma(t, ib - ia, ia);
}

// itp(f, to, rule, grid, I...)  interpolation.jl:63-80: the dims where location(f) != to are interpolated; knots
// (:53-56) are il/ir with loc = the field's location: f Center -> (I-e, I), f Vertex -> (I, I+e); the recursion
// (:19-25) applies the rule along the FIRST differing dim innermost and the LAST outermost.
template <class T>
OPR_HD T opr_itp(const OprFieldT<T>& f, const int* to, int nd, bool harmonic, const int* I) {
    int dims[3], m = 0;
    for (int d = 0; d < nd; ++d)
        if (f.loc[d] != to[d]) dims[m++] = d;
    if (m == 0) return opr_at(f, I);
    T v[8];
    for (int q = 0; q < (1 << m); ++q) {
        int J[3] = {I[0], I[1], I[2]};
        for (int b = 0; b < m; ++b) {
            const int d = dims[b];
            const int right = (q >> b) & 1;
            J[d] += (f.loc[d] == OPR_CENTER) ? right - 1 : right;
        }
        v[q] = opr_at(f, J);
    }
    for (int b = 0; b < m; ++b) {               // reduce the first differing dim first
        const int cnt = 1 << (m - 1 - b);
        for (int q = 0; q < cnt; ++q) v[q] = opr_rule(harmonic, (T)0.5, v[2 * q], v[2 * q + 1]);
    }
    return v[0];
}

// partial_derivatives.jl:14-21: (lerp(k, floc, Ir) * ∂(f, Ir) - lerp(k, floc, Il) * ∂(f, Il)) * iΔ, floc = flipped(loc(f), dim)
template <class T>
OPR_HD T opr_dkd(const OprFieldT<T>& f, const OprFieldT<T>& k, const T* id, int nd, int dim, const int* I) {
    int Ir[3], Il[3], floc[3] = {f.loc[0], f.loc[1], f.loc[2]};
    floc[dim] = 1 - floc[dim];
    opr_second_idx(f, dim, I, Ir, Il);
    const T a = opr_itp(k, floc, nd, false, Ir) * opr_partial(f, id, dim, Ir);
    const T b = opr_itp(k, floc, nd, false, Il) * opr_partial(f, id, dim, Il);
    return (a - b) * id[dim];
}

// One launch: dst[c][I] = OP(...)[I] for I in the launch box.
template <class T>
struct OprArgsT {
    int          oper, dim, nd;
    int          ndst;      // 1, or nd for GRAD / KGRAD
    OprFieldT<T> dst[3];
    OprFieldT<T> a[3];      // the source field f (a[0]) or the components of the vector field V
    OprFieldT<T> k;         // coefficient field of DKD / DIVG_GRAD / KGRAD
    T            id[3];     // inv_spacing
};
using OprArgs = OprArgsT<double>;

template <class T>
OPR_HD void opr_store(const OprFieldT<T>& f, const int* I, T v) {
    f.p[(long long)I[0] + (long long)I[1] * f.sy + (long long)I[2] * f.sz] = v;
}

template <class T>
OPR_HD void opr_apply(const OprArgsT<T>& g, int i, int j, int kk) {
    const int I[3] = {i, j, kk};
    switch (g.oper) {
    case OPR_LEFT: opr_store(g.dst[0], I, opr_left(g.a[0], g.dim, I)); break;
    case OPR_RIGHT: opr_store(g.dst[0], I, opr_right(g.a[0], g.dim, I)); break;
    case OPR_DELTA: opr_store(g.dst[0], I, opr_delta(g.a[0], g.dim, I)); break;
    case OPR_PARTIAL: opr_store(g.dst[0], I, opr_partial(g.a[0], g.id, g.dim, I)); break;
    case OPR_PARTIAL2: opr_store(g.dst[0], I, opr_partial2(g.a[0], g.id, g.dim, I)); break;
    case OPR_DKD: opr_store(g.dst[0], I, opr_dkd(g.a[0], g.k, g.id, g.nd, g.dim, I)); break;
    case OPR_LERP: opr_store(g.dst[0], I, opr_itp(g.a[0], g.dst[0].loc, g.nd, false, I)); break;
    case OPR_HLERP: opr_store(g.dst[0], I, opr_itp(g.a[0], g.dst[0].loc, g.nd, true, I)); break;
    case OPR_DIVG: {        // field_operators.jl:50-55: n-ary + folds left
        T s = opr_partial(g.a[0], g.id, 0, I);
        for (int d = 1; d < g.nd; ++d) s = s + opr_partial(g.a[d], g.id, d, I);
        opr_store(g.dst[0], I, s);
        break;
    }
    case OPR_LAPL: {        // field_operators.jl:72-77
        T s = opr_partial2(g.a[0], g.id, 0, I);
        for (int d = 1; d < g.nd; ++d) s = s + opr_partial2(g.a[0], g.id, d, I);
        opr_store(g.dst[0], I, s);
        break;
    }
    case OPR_DIVG_GRAD: {   // field_operators.jl:95-100
        T s = opr_dkd(g.a[0], g.k, g.id, g.nd, 0, I);
        for (int d = 1; d < g.nd; ++d) s = s + opr_dkd(g.a[0], g.k, g.id, g.nd, d, I);
        opr_store(g.dst[0], I, s);
        break;
    }
    case OPR_VMAG: {        // field_operators.jl:116-121: sqrt(sum_D lerp(V[D], Center())^2), x^2 = x*x
        const int ctr[3] = {OPR_CENTER, OPR_CENTER, OPR_CENTER};
        T c = opr_itp(g.a[0], ctr, g.nd, false, I);
        T s = c * c;
        for (int d = 1; d < g.nd; ++d) {
            c = opr_itp(g.a[d], ctr, g.nd, false, I);
            s = s + c * c;
        }
        opr_store(g.dst[0], I, sqrt(s));
        break;
    }
    case OPR_GRAD:          // test_grid_operators.jl:24-30: V.d[I] = ∂_d(C)
        for (int d = 0; d < g.nd; ++d) opr_store(g.dst[d], I, opr_partial(g.a[0], g.id, d, I));
        break;
    case OPR_KGRAD:         // test_grid_operators.jl:76-82: V.d[I] = lerp(χ, location(V.d)) * ∂_d(C)
        for (int d = 0; d < g.nd; ++d)
            opr_store(g.dst[d], I, opr_itp(g.k, g.dst[d].loc, g.nd, false, I) * opr_partial(g.a[0], g.id, d, I));
        break;
    default: break;
    }
}
