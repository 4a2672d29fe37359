// ops.cu -- the staggered-grid stencil kernels that `launch` runs (sm_100a).
//
// Arithmetic contract (DESIGN.md "Exact arithmetic"): every + - * / is one IEEE binary64 operation in the order the
// reference's Julia source evaluates it; this file is compiled with -fmad=false so nvcc never contracts a*b+c.
// fma() appears only where the reference writes muladd.  Formulas are the flattened forms of
//   examples/diffusion_2d.jl:8-19, examples/stokes_2d_inc_ve_T.jl:11-60, examples/stokes_3d_inc_ve_T.jl:11-77
// with the operator definitions of src/GridOperators/{GridOperators.jl:23-36,partial_derivatives.jl:2-5,
// field_operators.jl:50-59}:   Vertex along d: left=f[I], right=f[I+e_d];  Center along d: left=f[I-e_d], right=f[I];
// d_d f = (right-left)*inv_spacing_d;  n-ary + folds left.
//
// This translation unit holds the *generic* one-thread-per-cell kernels (any box, any layout); the tuned
// z-marching / vectorised kernels for the headline configs live in ops_fast.cu and are selected in chmy_run_op.
#include "common.cuh"
#include "operators.cuh"

// ---------------------------------------------------------------------------------------------- generic box kernel
template <class F>
__global__ void __launch_bounds__(256) k_box(const F f, const Box b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i < b.n[0] && j < b.n[1]) f(b.lo[0] + i, b.lo[1] + j, b.lo[2] + k);
}

template <class F>
static int launch_box(chmy_ctx* ctx, const F& f, const Box& b, cudaStream_t st) {
    if (b.n[0] <= 0 || b.n[1] <= 0 || b.n[2] <= 0) return CHMY_OK;
    const dim3 blk(64, b.n[1] > 1 ? 4 : 1, 1);
    const dim3 grd = grid_for(b, blk);
    k_box<F><<<grd, blk, 0, st>>>(f, b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

// Element type.  The reference's kernels are generic in the element type T of their fields (its suite runs Float32 and
// Float64, test/common.jl:9) but spell their constants as Float64 literals (0.5, 2.0, 3.0, 0.0).  With T = Float32 Julia
// promotes exactly the sub-expressions that meet such a literal to Float64 and rounds once when the result is stored into
// the Float32 array.  The functors below reproduce that operation by operation: values of type T are combined in T, `W`
// (double) marks the promoted sub-expressions, every store converts to T.  For T = double W == T and this is the plain
// binary64 arithmetic (pinned, for both element types, by tests/test_oracle_transliteration.py).
typedef double W;

// Julia Base.max / min with a Float64 zero: both arguments promoted (NaN-propagating, max(-0.0, +0.0) = +0.0)
__device__ __forceinline__ W jl_max0w(W v) { return (v != v) ? v : fmax(v, 0.0); }
__device__ __forceinline__ W jl_min0w(W v) { return (v != v) ? v : fmin(v, 0.0); }

// ---------------------------------------------------------------------------------------------- diffusion
// examples/diffusion_2d.jl:8-13   q.x=(V,C) q.y=(C,V) C=(C,C)
template <class T>
struct ComputeQ {
    FVT<T> qx, qy, C;
    T chi, idx, idy;
    __device__ void operator()(int i, int j, int k) const {
        const T c = fv_ld(C, i, j, k);
        fv_st(qx, i, j, k, (T)((-chi) * ((c - fv_ld(C, i - 1, j, k)) * idx)));
        fv_st(qy, i, j, k, (T)((-chi) * ((c - fv_ld(C, i, j - 1, k)) * idy)));
    }
};

// examples/diffusion_2d.jl:15-19   C -= dt * (dx(q.x) + dy(q.y))
template <class T>
struct UpdateC {
    FVT<T> C, qx, qy;
    T dt, idx, idy;
    __device__ void operator()(int i, int j, int k) const {
        const T dv = (fv_ld(qx, i + 1, j, k) - fv_ld(qx, i, j, k)) * idx +
                     (fv_ld(qy, i, j + 1, k) - fv_ld(qy, i, j, k)) * idy;
        fv_st(C, i, j, k, (T)(fv_ld(C, i, j, k) - dt * dv));
    }
};

// ---------------------------------------------------------------------------------------------- update_old
// stokes_3d_inc_ve_T.jl:11-21 / stokes_2d_inc_ve_T.jl:11-18 : same index I for every pair
template <class T, int NP>
struct UpdateOld {
    FVT<T> dst[NP], src[NP];
    __device__ void operator()(int i, int j, int k) const {
#pragma unroll
        for (int p = 0; p < NP; ++p) fv_st(dst[p], i, j, k, fv_ld(src[p], i, j, k));
    }
};

// ---------------------------------------------------------------------------------------------- stress
template <class T>
__device__ __forceinline__ W stress_res(T t, T to, W e2, T Gdt, T eta) {
    // r = -(tau - tau_old)/(G*dt) - tau/eta + 2.0*e      (stokes_3d_inc_ve_T.jl:34-39): the first two terms in T, the sum in W
    return (W)((-(t - to)) / Gdt - t / eta) + e2;
}

// stokes_2d_inc_ve_T.jl:20-34   tau.xy=(V,V), V.x=(V,C), V.y=(C,V)
template <class T>
struct Stress2 {
    FVT<T> txx, tyy, txy, Pr, dV, Vx, Vy, oxx, oyy, oxy;
    T idx, idy, eta, eta_ve, Gdt, dtau_Pr, dtau_r;
    __device__ void operator()(int i, int j, int k) const {
        const T vx = fv_ld(Vx, i, j, k), vy = fv_ld(Vy, i, j, k);
        const T exx = (fv_ld(Vx, i + 1, j, k) - vx) * idx;
        const T eyy = (fv_ld(Vy, i, j + 1, k) - vy) * idy;
        const W exy = 0.5 * (W)(T)((vx - fv_ld(Vx, i, j - 1, k)) * idy + (vy - fv_ld(Vy, i - 1, j, k)) * idx);
        const T dv  = exx + eyy;
        fv_st(dV, i, j, k, dv);
        fv_st(Pr, i, j, k, (T)(fv_ld(Pr, i, j, k) - (dv * eta_ve) * dtau_Pr));
        const W dv3 = (W)dv / 3.0;
        const T a = fv_ld(txx, i, j, k), b = fv_ld(tyy, i, j, k), c = fv_ld(txy, i, j, k);
        const W rxx = stress_res(a, fv_ld(oxx, i, j, k), 2.0 * ((W)exx - dv3), Gdt, eta);
        const W ryy = stress_res(b, fv_ld(oyy, i, j, k), 2.0 * ((W)eyy - dv3), Gdt, eta);
        const W rxy = stress_res(c, fv_ld(oxy, i, j, k), 2.0 * exy, Gdt, eta);
        fv_st(txx, i, j, k, (T)((W)a + (rxx * (W)eta_ve) * (W)dtau_r));
        fv_st(tyy, i, j, k, (T)((W)b + (ryy * (W)eta_ve) * (W)dtau_r));
        fv_st(txy, i, j, k, (T)((W)c + (rxy * (W)eta_ve) * (W)dtau_r));
    }
};

// stokes_3d_inc_ve_T.jl:23-46
template <class T>
struct Stress3 {
    FVT<T> t[6], Pr, dV, Vx, Vy, Vz, o[6];   // xx yy zz xy xz yz
    T idx, idy, idz, eta, eta_ve, Gdt, dtau_Pr, dtau_r;
    __device__ void operator()(int i, int j, int k) const {
        const T vx = fv_ld(Vx, i, j, k), vy = fv_ld(Vy, i, j, k), vz = fv_ld(Vz, i, j, k);
        const T exx = (fv_ld(Vx, i + 1, j, k) - vx) * idx;
        const T eyy = (fv_ld(Vy, i, j + 1, k) - vy) * idy;
        const T ezz = (fv_ld(Vz, i, j, k + 1) - vz) * idz;
        const W exy = 0.5 * (W)(T)((vx - fv_ld(Vx, i, j - 1, k)) * idy + (vy - fv_ld(Vy, i - 1, j, k)) * idx);
        const W exz = 0.5 * (W)(T)((vx - fv_ld(Vx, i, j, k - 1)) * idz + (vz - fv_ld(Vz, i - 1, j, k)) * idx);
        const W eyz = 0.5 * (W)(T)((vy - fv_ld(Vy, i, j, k - 1)) * idz + (vz - fv_ld(Vz, i, j - 1, k)) * idy);
        const T dv  = (exx + eyy) + ezz;
        fv_st(dV, i, j, k, dv);
        fv_st(Pr, i, j, k, (T)(fv_ld(Pr, i, j, k) - (dv * eta_ve) * dtau_Pr));
        const W dv3 = (W)dv / 3.0;
        const W e2[6] = {2.0 * ((W)exx - dv3), 2.0 * ((W)eyy - dv3), 2.0 * ((W)ezz - dv3), 2.0 * exy, 2.0 * exz, 2.0 * eyz};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const T a = fv_ld(t[c], i, j, k);
            const W r = stress_res(a, fv_ld(o[c], i, j, k), e2[c], Gdt, eta);
            fv_st(t[c], i, j, k, (T)((W)a + (r * (W)eta_ve) * (W)dtau_r));
        }
    }
};

// ---------------------------------------------------------------------------------------------- velocity
// stokes_2d_inc_ve_T.jl:36-43   (no literal: everything in T)
template <class T>
struct Velocity2 {
    FVT<T> Vx, Vy, rx, ry, Pr, txx, tyy, txy, rho;
    InclDevT<T> inc;
    T idx, idy, eta_ve, nudtau;
    __device__ void operator()(int i, int j, int k) const {
        const T p = fv_ld(Pr, i, j, k), sxy = fv_ld(txy, i, j, k);
        const T rvx = ((-((p - fv_ld(Pr, i - 1, j, k)) * idx)) + (fv_ld(txx, i, j, k) - fv_ld(txx, i - 1, j, k)) * idx) +
                      (fv_ld(txy, i, j + 1, k) - sxy) * idy;
        const T rg  = inc.active ? incl_eval(inc, i, j, k) : fv_ld(rho, i, j, k);
        const T rvy = (((-((p - fv_ld(Pr, i, j - 1, k)) * idy)) + (fv_ld(tyy, i, j, k) - fv_ld(tyy, i, j - 1, k)) * idy) +
                       (fv_ld(txy, i + 1, j, k) - sxy) * idx) - rg;
        fv_st(rx, i, j, k, rvx);
        fv_st(ry, i, j, k, rvy);
        fv_st(Vx, i, j, k, (T)(fv_ld(Vx, i, j, k) + (rvx * nudtau) / eta_ve));
        fv_st(Vy, i, j, k, (T)(fv_ld(Vy, i, j, k) + (rvy * nudtau) / eta_ve));
    }
};

// stokes_3d_inc_ve_T.jl:48-57
template <class T>
struct Velocity3 {
    FVT<T> Vx, Vy, Vz, rx, ry, rz, Pr, t[6], rho;   // t: xx yy zz xy xz yz
    InclDevT<T> inc;
    T idx, idy, idz, eta_ve, nudtau;
    __device__ void operator()(int i, int j, int k) const {
        const T p = fv_ld(Pr, i, j, k);
        const T sxy = fv_ld(t[3], i, j, k), sxz = fv_ld(t[4], i, j, k), syz = fv_ld(t[5], i, j, k);
        const T rvx = (((-((p - fv_ld(Pr, i - 1, j, k)) * idx)) + (fv_ld(t[0], i, j, k) - fv_ld(t[0], i - 1, j, k)) * idx) +
                       (fv_ld(t[3], i, j + 1, k) - sxy) * idy) + (fv_ld(t[4], i, j, k + 1) - sxz) * idz;
        const T rvy = (((-((p - fv_ld(Pr, i, j - 1, k)) * idy)) + (fv_ld(t[1], i, j, k) - fv_ld(t[1], i, j - 1, k)) * idy) +
                       (fv_ld(t[3], i + 1, j, k) - sxy) * idx) + (fv_ld(t[5], i, j, k + 1) - syz) * idz;
        const T rg  = inc.active ? incl_eval(inc, i, j, k) : fv_ld(rho, i, j, k);
        const T rvz = ((((-((p - fv_ld(Pr, i, j, k - 1)) * idz)) + (fv_ld(t[2], i, j, k) - fv_ld(t[2], i, j, k - 1)) * idz) +
                        (fv_ld(t[4], i + 1, j, k) - sxz) * idx) + (fv_ld(t[5], i, j + 1, k) - syz) * idy) - rg;
        fv_st(rx, i, j, k, rvx);
        fv_st(ry, i, j, k, rvy);
        fv_st(rz, i, j, k, rvz);
        fv_st(Vx, i, j, k, (T)(fv_ld(Vx, i, j, k) + (rvx * nudtau) / eta_ve));
        fv_st(Vy, i, j, k, (T)(fv_ld(Vy, i, j, k) + (rvy * nudtau) / eta_ve));
        fv_st(Vz, i, j, k, (T)(fv_ld(Vz, i, j, k) + (rvz * nudtau) / eta_ve));
    }
};

// ---------------------------------------------------------------------------------------------- thermal
// stokes_3d_inc_ve_T.jl:59-71 (2D: stokes_2d_inc_ve_T.jl:45-54):  -lam * d(T) in T; max(V, 0.0) / min(V, 0.0) and their
// products with left / right(T) in W; one rounding at the store
template <class T, int ND>
struct ThermalFlux {
    FVT<T> q[3], Tf, V[3];
    T lam, id[3];
    __device__ void operator()(int i, int j, int k) const {
        const T t = fv_ld(Tf, i, j, k);
        const T tm[3] = {fv_ld(Tf, i - 1, j, k), fv_ld(Tf, i, j - 1, k), ND > 2 ? fv_ld(Tf, i, j, k - 1) : (T)0.0};
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const W v = (W)fv_ld(V[d], i, j, k);
            fv_st(q[d], i, j, k, (T)(((W)(T)((-lam) * ((t - tm[d]) * id[d])) + jl_max0w(v) * (W)tm[d]) + jl_min0w(v) * (W)t));
        }
    }
};

// stokes_3d_inc_ve_T.jl:73-77
template <class T, int ND>
struct Thermal {
    FVT<T> Tf, To, q[3];
    T dt, id[3];
    __device__ void operator()(int i, int j, int k) const {
        T dv = (fv_ld(q[0], i + 1, j, k) - fv_ld(q[0], i, j, k)) * id[0] +
               (fv_ld(q[1], i, j + 1, k) - fv_ld(q[1], i, j, k)) * id[1];
        if (ND > 2) dv = dv + (fv_ld(q[2], i, j, k + 1) - fv_ld(q[2], i, j, k)) * id[2];
        fv_st(Tf, i, j, k, (T)(fv_ld(To, i, j, k) - dt * dv));
    }
};

// ---------------------------------------------------------------------------------------------- dispatch
static int expect_fields(const chmy_launch_desc* d, int nf, int ns, int allow_null_last) {
    CHMY_REQUIRE(d->nfields == nf, "op %d expects %d fields, got %d", d->op, nf, d->nfields);
    CHMY_REQUIRE(d->nscalars == ns, "op %d expects %d scalars, got %d", d->op, ns, d->nscalars);
    for (int i = 0; i < nf; ++i) {
        if (d->fields[i] == nullptr) {
            CHMY_REQUIRE(allow_null_last && i == nf - 1, "op %d: field %d is NULL", d->op, i);
            continue;
        }
        CHMY_REQUIRE(d->fields[i]->nd == d->grid.ndims, "op %d: field %d has %d dims, grid has %d", d->op, i,
                     d->fields[i]->nd, d->grid.ndims);
        CHMY_REQUIRE(d->fields[i]->dtype == d->fields[0]->dtype, "op %d: field %d has another element type than field 0", d->op, i);
    }
    return CHMY_OK;
}

static int expect_loc(const chmy_launch_desc* d, int idx, int lx, int ly, int lz) {
    const chmy_field* f = d->fields[idx];
    const int want[3] = {lx, ly, lz};
    for (int a = 0; a < d->grid.ndims; ++a) {
        CHMY_REQUIRE(f->loc[a] == want[a], "op %d: field %d has the wrong staggered location along dim %d", d->op, idx, a + 1);
        CHMY_REQUIRE(f->d[a] == d->grid.n[a] + (want[a] == CHMY_VERTEX ? 1 : 0),
                     "op %d: field %d size %lld along dim %d does not match the grid", d->op, idx, f->d[a], a + 1);
    }
    return CHMY_OK;
}

enum { C_ = CHMY_CENTER, V_ = CHMY_VERTEX };

// location tables: tensor components xx yy [zz] xy [xz yz]; vector components
static const int TLOC3[6][3] = {{C_, C_, C_}, {C_, C_, C_}, {C_, C_, C_}, {V_, V_, C_}, {V_, C_, V_}, {C_, V_, V_}};
static const int TLOC2[3][3] = {{C_, C_, C_}, {C_, C_, C_}, {V_, V_, C_}};
static const int VLOC[3][3]  = {{V_, C_, C_}, {C_, V_, C_}, {C_, C_, V_}};

static int expect_tensor(const chmy_launch_desc* d, int first, int nd) {
    const int nt = nd == 2 ? 3 : 6;
    for (int c = 0; c < nt; ++c) {
        const int* L = nd == 2 ? TLOC2[c] : TLOC3[c];
        CHMY_TRY(expect_loc(d, first + c, L[0], L[1], L[2]));
    }
    return CHMY_OK;
}
static int expect_vector(const chmy_launch_desc* d, int first, int nd) {
    for (int c = 0; c < nd; ++c) CHMY_TRY(expect_loc(d, first + c, VLOC[c][0], VLOC[c][1], VLOC[c][2]));
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- grid operators
// CHMY_OP_OPERATOR (operators.cuh): every field must sit where the reference's operator reads / produces it.
static int expect_size(const chmy_launch_desc* d, int idx) {
    const chmy_field* f = d->fields[idx];
    for (int a = 0; a < d->grid.ndims; ++a)
        CHMY_REQUIRE(f->d[a] == d->grid.n[a] + (f->loc[a] == CHMY_VERTEX ? 1 : 0),
                     "operator %d: field %d size %lld along dim %d does not match the grid", d->oper, idx, f->d[a], a + 1);
    return CHMY_OK;
}
static bool loc_is(const chmy_field* f, const chmy_field* src, int nd, int flip_dim) {   // loc(f) == flipped(loc(src), flip_dim)
    for (int a = 0; a < nd; ++a)
        if (f->loc[a] != (a == flip_dim ? 1 - src->loc[a] : src->loc[a])) return false;
    return true;
}

static int validate_operator(const chmy_launch_desc* d) {
    const int nd = d->grid.ndims, op = d->oper;
    CHMY_REQUIRE(op >= CHMY_OPER_LEFT && op <= CHMY_OPER_KGRAD, "unknown operator id %d", op);
    CHMY_REQUIRE(d->nscalars == 0, "operator %d takes no scalars", op);
    int nf = 2;
    if (op == CHMY_OPER_DKD || op == CHMY_OPER_DIVG_GRAD) nf = 3;
    if (op == CHMY_OPER_DIVG || op == CHMY_OPER_VMAG || op == CHMY_OPER_GRAD) nf = nd + 1;
    if (op == CHMY_OPER_KGRAD) nf = nd + 2;
    CHMY_REQUIRE(d->nfields == nf, "operator %d expects %d fields on a %dD grid, got %d", op, nf, nd, d->nfields);
    for (int i = 0; i < nf; ++i) {
        CHMY_REQUIRE(d->fields[i] != nullptr, "operator %d: field %d is NULL", op, i);
        CHMY_REQUIRE(d->fields[i]->nd == nd, "operator %d: field %d has %d dims, grid has %d", op, i, d->fields[i]->nd, nd);
        CHMY_REQUIRE(d->fields[i]->dtype == d->fields[0]->dtype, "operator %d: the fields must share one element type", op);
        CHMY_TRY(expect_size(d, i));
    }
    chmy_field* const* F = d->fields;
    const int nout = (op == CHMY_OPER_GRAD || op == CHMY_OPER_KGRAD) ? nd : 1;
    for (int o = 0; o < nout; ++o)
        for (int i = nout; i < nf; ++i)
            CHMY_REQUIRE(F[o] != F[i], "operator %d: the destination must not alias a source (neighbouring points are read)", op);
    if (op <= CHMY_OPER_DKD) CHMY_REQUIRE(d->oper_dim >= 0 && d->oper_dim < nd, "operator %d: dim %d out of range", op, d->oper_dim + 1);
    switch (op) {
    case CHMY_OPER_LEFT: case CHMY_OPER_RIGHT: case CHMY_OPER_DELTA: case CHMY_OPER_PARTIAL:
        CHMY_REQUIRE(loc_is(F[0], F[1], nd, d->oper_dim), "operator %d: dst must be located at flipped(location(f), dim)", op);
        break;
    case CHMY_OPER_PARTIAL2: case CHMY_OPER_DKD: case CHMY_OPER_LAPL: case CHMY_OPER_DIVG_GRAD:
        CHMY_REQUIRE(loc_is(F[0], F[1], nd, -1), "operator %d: dst must be located at location(f)", op);
        break;
    case CHMY_OPER_DIVG:
        for (int c = 0; c < nd; ++c)
            CHMY_REQUIRE(loc_is(F[0], F[1 + c], nd, c), "divg: dst must be located at flipped(location(V.%c), %d)", "xyz"[c], c + 1);
        break;
    case CHMY_OPER_VMAG:
        for (int a = 0; a < nd; ++a) CHMY_REQUIRE(F[0]->loc[a] == CHMY_CENTER, "vmag: dst must be located at Center()");
        break;
    case CHMY_OPER_GRAD: case CHMY_OPER_KGRAD:
        for (int c = 0; c < nd; ++c)
            CHMY_REQUIRE(loc_is(F[c], F[nd], nd, c), "operator %d: V.%c must be located at flipped(location(f), %d)", op, "xyz"[c], c + 1);
        break;
    default: break;   // LERP / HLERP: any two locations
    }
    return CHMY_OK;
}

template <class T>
struct OperatorF {
    OprArgsT<T> g;
    __device__ void operator()(int i, int j, int k) const { opr_apply(g, i, j, k); }
};

template <class T>
static OprFieldT<T> opr_view(const chmy_field* f) {
    OprFieldT<T> v;
    v.p = reinterpret_cast<T*>(f->p0); v.sy = f->nd > 1 ? f->stride[1] : 0; v.sz = f->nd > 2 ? f->stride[2] : 0;
    for (int a = 0; a < 3; ++a) v.loc[a] = f->loc[a];
    return v;
}

template <class T>
static int run_operator_t(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st) {
    const int nd = d->grid.ndims, op = d->oper;
    chmy_field* const* F = d->fields;
    OperatorF<T> f;
    memset(&f, 0, sizeof(f));
    f.g.oper = op; f.g.dim = d->oper_dim; f.g.nd = nd;
    for (int a = 0; a < 3; ++a) f.g.id[a] = a < nd ? (T)d->grid.inv_spacing[a] : (T)0.0;
    const int nout = (op == CHMY_OPER_GRAD || op == CHMY_OPER_KGRAD) ? nd : 1;
    f.g.ndst = nout;
    for (int o = 0; o < nout; ++o) f.g.dst[o] = opr_view<T>(F[o]);
    const bool vec = op == CHMY_OPER_DIVG || op == CHMY_OPER_VMAG;
    for (int c = 0; c < (vec ? nd : 1); ++c) f.g.a[c] = opr_view<T>(F[nout + c]);
    if (op == CHMY_OPER_DKD || op == CHMY_OPER_DIVG_GRAD || op == CHMY_OPER_KGRAD) f.g.k = opr_view<T>(F[nout + 1]);
    for (int o = 0; o < nout; ++o) F[o]->frame_dirty(0);   // not a ping-pong op: a shadow copy of dst goes stale
    return launch_box(ctx, f, box, st);
}

static int run_operator(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st) {
    if (d->fields[0]->dtype == CHMY_F32) return run_operator_t<float>(ctx, d, box, st);
    return run_operator_t<double>(ctx, d, box, st);
}

int chmy_validate_op(const chmy_launch_desc* d) {
    const int nd = d->grid.ndims;
    const int nt = nd == 2 ? 3 : 6;
    switch (d->op) {
    case CHMY_OP_NONE: return CHMY_OK;
    case CHMY_OP_COMPUTE_Q:
        CHMY_REQUIRE(nd == 2, "compute_q! is defined for 2D grids");
        CHMY_TRY(expect_fields(d, 3, 1, 0));
        CHMY_TRY(expect_vector(d, 0, 2));
        return expect_loc(d, 2, C_, C_, C_);
    case CHMY_OP_UPDATE_C:
        CHMY_REQUIRE(nd == 2, "update_C! is defined for 2D grids");
        CHMY_TRY(expect_fields(d, 3, 1, 0));
        CHMY_TRY(expect_loc(d, 0, C_, C_, C_));
        return expect_vector(d, 1, 2);
    case CHMY_OP_UPDATE_OLD:
        CHMY_REQUIRE(nd == 2 || nd == 3, "update_old! is defined for 2D and 3D grids");
        CHMY_TRY(expect_fields(d, 2 * (1 + nt), 0, 0));
        CHMY_TRY(expect_loc(d, 0, C_, C_, C_));
        CHMY_TRY(expect_tensor(d, 1, nd));
        CHMY_TRY(expect_loc(d, 1 + nt, C_, C_, C_));
        return expect_tensor(d, 2 + nt, nd);
    case CHMY_OP_UPDATE_STRESS:
        CHMY_REQUIRE(nd == 2 || nd == 3, "update_stress! is defined for 2D and 3D grids");
        CHMY_TRY(expect_fields(d, 2 * nt + 2 + nd, 6, 0));
        CHMY_TRY(expect_tensor(d, 0, nd));
        CHMY_TRY(expect_loc(d, nt, C_, C_, C_));
        CHMY_TRY(expect_loc(d, nt + 1, C_, C_, C_));
        CHMY_TRY(expect_vector(d, nt + 2, nd));
        return expect_tensor(d, nt + 2 + nd, nd);
    case CHMY_OP_UPDATE_VELOCITY:
        CHMY_REQUIRE(nd == 2 || nd == 3, "update_velocity! is defined for 2D and 3D grids");
        CHMY_TRY(expect_fields(d, 2 * nd + 1 + nt + 1, 2, 1));
        CHMY_TRY(expect_vector(d, 0, nd));
        CHMY_TRY(expect_vector(d, nd, nd));
        CHMY_TRY(expect_loc(d, 2 * nd, C_, C_, C_));
        CHMY_TRY(expect_tensor(d, 2 * nd + 1, nd));
        if (d->fields[2 * nd + 1 + nt] != nullptr) {
            CHMY_REQUIRE(!d->rho_g.active, "update_velocity!: give rho_g either as a field or as a FunctionField");
            return expect_loc(d, 2 * nd + 1 + nt, VLOC[nd - 1][0], VLOC[nd - 1][1], VLOC[nd - 1][2]);
        }
        CHMY_REQUIRE(d->rho_g.active, "update_velocity!: rho_g missing");
        for (int a = 0; a < nd; ++a)
            CHMY_REQUIRE(d->rho_g.loc[a] == VLOC[nd - 1][a], "update_velocity!: rho_g FunctionField location mismatch");
        return CHMY_OK;
    case CHMY_OP_UPDATE_THERMAL_FLUX:
        CHMY_REQUIRE(nd == 2 || nd == 3, "update_thermal_flux! is defined for 2D and 3D grids");
        CHMY_TRY(expect_fields(d, 2 * nd + 1, 1, 0));
        CHMY_TRY(expect_vector(d, 0, nd));
        CHMY_TRY(expect_loc(d, nd, C_, C_, C_));
        return expect_vector(d, nd + 1, nd);
    case CHMY_OP_UPDATE_THERMAL:
        CHMY_REQUIRE(nd == 2 || nd == 3, "update_thermal! is defined for 2D and 3D grids");
        CHMY_TRY(expect_fields(d, 2 + nd, 1, 0));
        CHMY_TRY(expect_loc(d, 0, C_, C_, C_));
        CHMY_TRY(expect_loc(d, 1, C_, C_, C_));
        return expect_vector(d, 2, nd);
    case CHMY_OP_OPERATOR: return validate_operator(d);
    default: chmy_set_error("unknown op id %d", d->op); return CHMY_ERR_ARG;
    }
}

template <class T>
static InclDevT<T> make_incl(const chmy_launch_desc* d) {
    InclDevT<T> q;
    memset(&q, 0, sizeof(q));
    q.active = d->rho_g.active;
    q.nd     = d->grid.ndims;
    for (int a = 0; a < 3; ++a) {
        q.loc[a]     = d->rho_g.loc[a];
        q.origin[a]  = (T)d->grid.origin[a];
        q.spacing[a] = (T)d->grid.spacing[a];
        q.c0[a]      = (T)d->rho_g.c0[a];
    }
    q.r2  = (T)d->rho_g.r * (T)d->rho_g.r;   // r^2 -> r*r (Base.literal_pow), in the element type
    q.in  = (T)d->rho_g.in;
    q.out = (T)d->rho_g.out;
    return q;
}

// scalars and grid numbers cross the ABI as doubles holding values of the element type (a Float32 widens exactly)
template <class T>
static int run_op_generic_t(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st) {
    const int nd = d->grid.ndims;
    const int nt = nd == 2 ? 3 : 6;
    T id[3], s[CHMY_MAX_SCALARS];
    for (int a = 0; a < 3; ++a) id[a] = (T)d->grid.inv_spacing[a];
    for (int q = 0; q < CHMY_MAX_SCALARS; ++q) s[q] = q < d->nscalars ? (T)d->scalars[q] : (T)0;
    chmy_field* const* F = d->fields;
    auto V = [](const chmy_field* f) { return f->viewT<T>(); };
    switch (d->op) {
    case CHMY_OP_COMPUTE_Q: {
        ComputeQ<T> f{V(F[0]), V(F[1]), V(F[2]), s[0], id[0], id[1]};
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_C: {
        UpdateC<T> f{V(F[0]), V(F[1]), V(F[2]), s[0], id[0], id[1]};
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_OLD: {
        if (nd == 2) {
            UpdateOld<T, 4> f;
            for (int p = 0; p < 4; ++p) { f.src[p] = V(F[p]); f.dst[p] = V(F[4 + p]); }
            return launch_box(ctx, f, box, st);
        }
        UpdateOld<T, 7> f;
        for (int p = 0; p < 7; ++p) { f.src[p] = V(F[p]); f.dst[p] = V(F[7 + p]); }
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_STRESS: {
        const T Gdt = s[2] * s[3];   // G*dt in the element type; evaluated once per cell in the reference, same value
        if (nd == 2) {
            Stress2<T> f{V(F[0]), V(F[1]), V(F[2]), V(F[3]), V(F[4]), V(F[5]), V(F[6]),
                         V(F[7]), V(F[8]), V(F[9]), id[0], id[1], s[0], s[1], Gdt, s[4], s[5]};
            return launch_box(ctx, f, box, st);
        }
        Stress3<T> f;
        for (int c = 0; c < 6; ++c) { f.t[c] = V(F[c]); f.o[c] = V(F[11 + c]); }
        f.Pr = V(F[6]); f.dV = V(F[7]);
        f.Vx = V(F[8]); f.Vy = V(F[9]); f.Vz = V(F[10]);
        f.idx = id[0]; f.idy = id[1]; f.idz = id[2];
        f.eta = s[0]; f.eta_ve = s[1]; f.Gdt = Gdt; f.dtau_Pr = s[4]; f.dtau_r = s[5];
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_VELOCITY: {
        const chmy_field* rho = F[2 * nd + 1 + nt];
        const FVT<T> rv = rho ? V(rho) : FVT<T>{nullptr, 0, 0};
        if (nd == 2) {
            Velocity2<T> f{V(F[0]), V(F[1]), V(F[2]), V(F[3]), V(F[4]), V(F[5]),
                           V(F[6]), V(F[7]), rv, make_incl<T>(d), id[0], id[1], s[0], s[1]};
            return launch_box(ctx, f, box, st);
        }
        Velocity3<T> f;
        f.Vx = V(F[0]); f.Vy = V(F[1]); f.Vz = V(F[2]);
        f.rx = V(F[3]); f.ry = V(F[4]); f.rz = V(F[5]);
        f.Pr = V(F[6]);
        for (int c = 0; c < 6; ++c) f.t[c] = V(F[7 + c]);
        f.rho = rv; f.inc = make_incl<T>(d);
        f.idx = id[0]; f.idy = id[1]; f.idz = id[2]; f.eta_ve = s[0]; f.nudtau = s[1];
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_THERMAL_FLUX: {
        if (nd == 2) {
            ThermalFlux<T, 2> f;
            for (int c = 0; c < 2; ++c) { f.q[c] = V(F[c]); f.V[c] = V(F[3 + c]); f.id[c] = id[c]; }
            f.Tf = V(F[2]); f.lam = s[0];
            return launch_box(ctx, f, box, st);
        }
        ThermalFlux<T, 3> f;
        for (int c = 0; c < 3; ++c) { f.q[c] = V(F[c]); f.V[c] = V(F[4 + c]); f.id[c] = id[c]; }
        f.Tf = V(F[3]); f.lam = s[0];
        return launch_box(ctx, f, box, st);
    }
    case CHMY_OP_UPDATE_THERMAL: {
        if (nd == 2) {
            Thermal<T, 2> f;
            f.Tf = V(F[0]); f.To = V(F[1]);
            for (int c = 0; c < 2; ++c) { f.q[c] = V(F[2 + c]); f.id[c] = id[c]; }
            f.dt = s[0];
            return launch_box(ctx, f, box, st);
        }
        Thermal<T, 3> f;
        f.Tf = V(F[0]); f.To = V(F[1]);
        for (int c = 0; c < 3; ++c) { f.q[c] = V(F[2 + c]); f.id[c] = id[c]; }
        f.dt = s[0];
        return launch_box(ctx, f, box, st);
    }
    default: chmy_set_error("unknown op id %d", d->op); return CHMY_ERR_ARG;
    }
}

int chmy_run_op_generic(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st) {
    if (d->op == CHMY_OP_OPERATOR) return run_operator(ctx, d, box, st);
    const chmy_field* f0 = d->nfields > 0 ? d->fields[0] : nullptr;
    if (f0 && f0->dtype == CHMY_F32) return run_op_generic_t<float>(ctx, d, box, st);
    return run_op_generic_t<double>(ctx, d, box, st);
}
