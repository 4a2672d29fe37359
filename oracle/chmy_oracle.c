/*
 * chmy_oracle.c -- CPU restatement of the Chmy.jl hot path.  TEST INFRASTRUCTURE ONLY.
 * See chmy_oracle.h for scope, pinning status and conventions.
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
 */
#include "chmy_oracle.h"
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* math in the element type: fma/sqrt/fabs must be the binary32 functions in the Float32 build (an fma evaluated in
 * binary64 and rounded again is not fmaf) */
#ifdef OG_F32
#define FMA(a, b, c) fmaf((og_real)(a), (og_real)(b), (og_real)(c))
#define SQRT(x) sqrtf(x)
#define FABS(x) fabsf(x)
#else
#define FMA(a, b, c) fma(a, b, c)
#define SQRT(x) sqrt(x)
#define FABS(x) fabs(x)
#endif

int og_real_bytes(void) { return (int)sizeof(og_real); }

#define IDX(f, i, j, k)                                                                       \
    ((size_t)((i) + (f)->o[0]) +                                                              \
     (size_t)(f)->sd[0] * ((size_t)((j) + (f)->o[1]) + (size_t)(f)->sd[1] * (size_t)((k) + (f)->o[2])))
#define AT(f, i, j, k) ((f)->data[IDX(f, i, j, k)])

int og_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* a launcher may have exported OMP_NUM_THREADS=1 for its workers (torchrun does): the CPU arm asks for the cores itself */
void og_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------- grid ------------------- */

/* src/Grids/uniform_axis.jl:7-11 : spacing = extent/len ; inv_spacing = inv(spacing) */
void og_grid_init(og_grid* g, int nd, const int64_t* n, const og_real* origin, const og_real* extent) {
    memset(g, 0, sizeof(*g));
    g->nd = nd;
    for (int d = 0; d < 3; ++d) {
        if (d < nd) {
            g->n[d]           = n[d];
            g->origin[d]      = origin[d];
            g->extent[d]      = extent[d];
            g->spacing[d]     = extent[d] / (og_real)n[d];
            g->inv_spacing[d] = (og_real)1.0 / g->spacing[d];
        } else {
            g->n[d] = 1;
            g->spacing[d] = g->inv_spacing[d] = 0.0;
        }
    }
}

/* src/Grids/uniform_axis.jl:18-19
 *   vertex(ax,i) = muladd(i-1, spacing, origin)
 *   center(ax,i) = muladd(i-1, spacing, muladd(0.5, spacing, origin)) */
og_real og_coord(const og_grid* g, int dim, int loc, int64_t i) {
    og_real im1 = (og_real)(i - 1);
    if (loc == OG_VERTEX) return FMA(im1, g->spacing[dim], g->origin[dim]);
    return FMA(im1, g->spacing[dim], FMA(0.5, g->spacing[dim], g->origin[dim]));
}

/* src/Fields/field.jl:56-62 : dims = size(grid,loc) ; data_size = dims + 4*halo (halo = 1) */
int64_t og_field_storage_len(const og_grid* g, const int32_t* loc) {
    int64_t len = 1;
    for (int d = 0; d < g->nd; ++d) len *= g->n[d] + (loc[d] == OG_VERTEX ? 1 : 0) + 4;
    return len;
}

void og_field_init(og_field* f, const og_grid* g, const int32_t* loc, og_real* data) {
    memset(f, 0, sizeof(*f));
    f->nd = g->nd;
    for (int d = 0; d < 3; ++d) {
        if (d < g->nd) {
            f->loc[d] = loc[d];
            f->d[d]   = g->n[d] + (loc[d] == OG_VERTEX ? 1 : 0);   /* abstract_axis.jl:47-50 */
            f->sd[d]  = f->d[d] + 4;
            f->o[d]   = 1;                                         /* field.jl:18: data[I + 2H], 1-based */
        } else {
            f->loc[d] = OG_CENTER;
            f->d[d] = 1; f->sd[d] = 1; f->o[d] = 0;
        }
    }
    f->data = data;
}

/* ---------------------------------------------------------------- operators --------------- */

/* src/GridOperators/GridOperators.jl:23-36 + field_operators.jl:2-13 (from = flipped(loc, dim)):
 *   Vertex along dim : left = f[I],       right = f[I+e]
 *   Center along dim : left = f[I-e],     right = f[I]                                        */
static inline og_real f_left(const og_field* f, int dim, int64_t i, int64_t j, int64_t k) {
    int64_t I[3] = {i, j, k};
    if (f->loc[dim] == OG_CENTER) I[dim] -= 1;
    return AT(f, I[0], I[1], I[2]);
}
static inline og_real f_right(const og_field* f, int dim, int64_t i, int64_t j, int64_t k) {
    int64_t I[3] = {i, j, k};
    if (f->loc[dim] == OG_VERTEX) I[dim] += 1;
    return AT(f, I[0], I[1], I[2]);
}
/* partial_derivatives.jl:2,5 : delta = right - left ; d = delta * inv_spacing */
static inline og_real f_d(const og_grid* g, const og_field* f, int dim, int64_t i, int64_t j, int64_t k) {
    return (f_right(f, dim, i, j, k) - f_left(f, dim, i, j, k)) * g->inv_spacing[dim];
}
og_real og_partial(const og_grid* g, const og_field* f, int dim, int64_t i, int64_t j, int64_t k) {
    return f_d(g, f, dim, i, j, k);
}

/* partial_derivatives.jl:7-12 with field_operators.jl:26-30 (from = loc):
 * Ir = ir(flip(L), L) : L=Center -> I+e ; L=Vertex -> I ;  Il = il(flip(L), L): Center -> I ; Vertex -> I-e */
static inline void second_idx(const og_field* f, int dim, int64_t* Ir, int64_t* Il) {
    if (f->loc[dim] == OG_CENTER) Ir[dim] += 1; else Il[dim] -= 1;
}
og_real og_partial2(const og_grid* g, const og_field* f, int dim, int64_t i, int64_t j, int64_t k) {
    int64_t Ir[3] = {i, j, k}, Il[3] = {i, j, k};
    second_idx(f, dim, Ir, Il);
    return (f_d(g, f, dim, Ir[0], Ir[1], Ir[2]) - f_d(g, f, dim, Il[0], Il[1], Il[2])) * g->inv_spacing[dim];
}

/* interpolation.jl:14 (Linear rule = muladd(t, b-a, a)), :19-25 (recursion: last differing dim outermost),
 * :29-33 (uniform weights 0.5), :53-56 (knots: il/ir with loc = field location, from = target location) */
static inline og_real itp_rule(int harmonic, og_real t, og_real a, og_real b) {
    if (!harmonic) return FMA(t, b - a, a);                       /* Linear:         muladd(t, b - a, a)              */
    const og_real ia = (og_real)1.0 / a, ib = (og_real)1.0 / b;   /* HarmonicLinear: inv(muladd(t, inv(b)-inv(a), inv(a))) */
    return (og_real)1.0 / FMA(t, ib - ia, ia);
}
static og_real itp_rec(const og_field* f, const int32_t* to, int top, int64_t* I, int harmonic) {
    int d = top;
    while (d >= 0 && f->loc[d] == to[d]) --d;
    if (d < 0) return AT(f, I[0], I[1], I[2]);
    int64_t save = I[d];
    /* field Center -> target Vertex: (I-1, I) ; field Vertex -> target Center: (I, I+1) */
    int64_t il = (f->loc[d] == OG_CENTER) ? save - 1 : save;
    int64_t ir = (f->loc[d] == OG_CENTER) ? save : save + 1;
    I[d] = il; og_real a = itp_rec(f, to, d - 1, I, harmonic);
    I[d] = ir; og_real b = itp_rec(f, to, d - 1, I, harmonic);
    I[d] = save;
    return itp_rule(harmonic, 0.5, a, b);
}
static og_real og_itp(const og_grid* g, const og_field* f, const int32_t* to, int harmonic, int64_t i, int64_t j, int64_t k) {
    int64_t I[3] = {i, j, k};
    int32_t to3[3] = {0, 0, 0};
    for (int d = 0; d < g->nd; ++d) to3[d] = to[d];
    for (int d = g->nd; d < 3; ++d) to3[d] = f->loc[d];
    return itp_rec(f, to3, g->nd - 1, I, harmonic);
}
og_real og_lerp(const og_grid* g, const og_field* f, const int32_t* to, int64_t i, int64_t j, int64_t k) {
    return og_itp(g, f, to, 0, i, j, k);
}
/* interpolation.jl:15,94 */
og_real og_hlerp(const og_grid* g, const og_field* f, const int32_t* to, int64_t i, int64_t j, int64_t k) {
    return og_itp(g, f, to, 1, i, j, k);
}

/* partial_derivatives.jl:14-21 : (lerp(k,floc,Ir)*d(f,Ir) - lerp(k,floc,Il)*d(f,Il)) * inv_spacing */
og_real og_dkd(const og_grid* g, const og_field* f, const og_field* kf, int dim, int64_t i, int64_t j, int64_t k) {
    int32_t floc[3] = {f->loc[0], f->loc[1], f->loc[2]};
    floc[dim] = 1 - floc[dim];
    int64_t Ir[3] = {i, j, k}, Il[3] = {i, j, k};
    second_idx(f, dim, Ir, Il);
    og_real a = og_lerp(g, kf, floc, Ir[0], Ir[1], Ir[2]) * f_d(g, f, dim, Ir[0], Ir[1], Ir[2]);
    og_real b = og_lerp(g, kf, floc, Il[0], Il[1], Il[2]) * f_d(g, f, dim, Il[0], Il[1], Il[2]);
    return (a - b) * g->inv_spacing[dim];
}

/* Field-level application of one operator over a box: dst[I] = OP(src...)[I] (test/test_grid_operators.jl:21-126 run
 * such bodies over I in [0, n+1]^N).  kind: 1 left 2 right 3 delta 4 partial 5 partial2 6 dkd 7 lerp 8 hlerp 9 divg
 * 10 lapl 11 divg_grad 12 vmag 13 grad (dst[d] = d_d f) 14 kgrad (dst[d] = lerp(k, loc(dst[d])) * d_d f).
 * src: f (or the nd vector components for divg / vmag); sums fold left (field_operators.jl:50-121). */
void og_apply_operator(const og_grid* g, int kind, int dim, og_field* const* dst, const og_field* const* src,
                       const og_field* kf, const int64_t* lo, const int64_t* hi) {
    const int nd = g->nd;
    const int32_t ctr[3] = {OG_CENTER, OG_CENTER, OG_CENTER};
    int64_t l3[3] = {0, 0, 0}, h3[3] = {0, 0, 0};
    for (int d = 0; d < nd; ++d) { l3[d] = lo[d]; h3[d] = hi[d]; }
    for (int64_t k = l3[2]; k <= h3[2]; ++k)
        for (int64_t j = l3[1]; j <= h3[1]; ++j)
            for (int64_t i = l3[0]; i <= h3[0]; ++i) {
                og_real s = 0.0;
                switch (kind) {
                case 1: AT(dst[0], i, j, k) = f_left(src[0], dim, i, j, k); break;
                case 2: AT(dst[0], i, j, k) = f_right(src[0], dim, i, j, k); break;
                case 3: AT(dst[0], i, j, k) = f_right(src[0], dim, i, j, k) - f_left(src[0], dim, i, j, k); break;
                case 4: AT(dst[0], i, j, k) = og_partial(g, src[0], dim, i, j, k); break;
                case 5: AT(dst[0], i, j, k) = og_partial2(g, src[0], dim, i, j, k); break;
                case 6: AT(dst[0], i, j, k) = og_dkd(g, src[0], kf, dim, i, j, k); break;
                case 7: AT(dst[0], i, j, k) = og_lerp(g, src[0], dst[0]->loc, i, j, k); break;
                case 8: AT(dst[0], i, j, k) = og_hlerp(g, src[0], dst[0]->loc, i, j, k); break;
                case 9:
                    s = og_partial(g, src[0], 0, i, j, k);
                    for (int d = 1; d < nd; ++d) s = s + og_partial(g, src[d], d, i, j, k);
                    AT(dst[0], i, j, k) = s; break;
                case 10:
                    s = og_partial2(g, src[0], 0, i, j, k);
                    for (int d = 1; d < nd; ++d) s = s + og_partial2(g, src[0], d, i, j, k);
                    AT(dst[0], i, j, k) = s; break;
                case 11:
                    s = og_dkd(g, src[0], kf, 0, i, j, k);
                    for (int d = 1; d < nd; ++d) s = s + og_dkd(g, src[0], kf, d, i, j, k);
                    AT(dst[0], i, j, k) = s; break;
                case 12: {
                    og_real c = og_lerp(g, src[0], ctr, i, j, k);
                    s = c * c;
                    for (int d = 1; d < nd; ++d) { c = og_lerp(g, src[d], ctr, i, j, k); s = s + c * c; }
                    AT(dst[0], i, j, k) = SQRT(s); break;
                }
                case 13:
                    for (int d = 0; d < nd; ++d) AT(dst[d], i, j, k) = og_partial(g, src[0], d, i, j, k);
                    break;
                case 14:
                    for (int d = 0; d < nd; ++d)
                        AT(dst[d], i, j, k) = og_lerp(g, kf, dst[d]->loc, i, j, k) * og_partial(g, src[0], d, i, j, k);
                    break;
                default: break;
                }
            }
}

/* Julia Base max/min on Float64: NaN if either is NaN; max(-0.0,+0.0) = +0.0, min = -0.0 */
static inline og_real jl_max(og_real a, og_real b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static inline og_real jl_min(og_real a, og_real b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}
/* the same on promoted operands: max(x::Float32, 0.0) is max(Float64(x), 0.0) in Julia */
static inline og_wide jl_maxw(og_wide a, og_wide b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static inline og_wide jl_minw(og_wide a, og_wide b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}

/* ---------------------------------------------------------------- set / reduce ------------ */

/* src/Fields/field.jl:121-124,131-142 (_set_continuous! over the interior) with the init_incl body
 * examples/stokes_3d_inc_ve_T.jl:125 : ifelse((x-x0)^2 + (y-y0)^2 + (z-z0)^2 < r^2, in, out)
 * (literal ^2 -> x*x ; n-ary + folds left) */
static inline og_real incl_value(const og_grid* g, const og_inclusion* inc, const int32_t* loc,
                                int64_t i, int64_t j, int64_t k) {
    int64_t I[3] = {i, j, k};
    og_real s = 0.0;
    for (int d = 0; d < g->nd; ++d) {
        og_real c = og_coord(g, d, loc[d], I[d]) - inc->c0[d];
        og_real c2 = c * c;
        s = (d == 0) ? c2 : s + c2;
    }
    return (s < inc->r * inc->r) ? inc->in : inc->out;
}

void og_set_inclusion(const og_grid* g, og_field* f, const og_inclusion* inc) {
    int64_t k0 = f->nd > 2 ? 1 : 0, k1 = f->nd > 2 ? f->d[2] : 0;
    int64_t j0 = f->nd > 1 ? 1 : 0, j1 = f->nd > 1 ? f->d[1] : 0;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int64_t k = k0; k <= k1; ++k)
        for (int64_t j = j0; j <= j1; ++j)
            for (int64_t i = 1; i <= f->d[0]; ++i)
                AT(f, i, j, k) = incl_value(g, inc, f->loc, i, j, k);
}

/* driver code: maximum(abs.(interior(f)))  e.g. examples/stokes_3d_inc_ve_T.jl:158,172-175 */
og_real og_maxabs_interior(const og_field* f) {
    og_real m = 0.0;
    int nanflag = 0;
    int64_t k0 = f->nd > 2 ? 1 : 0, k1 = f->nd > 2 ? f->d[2] : 0;
    int64_t j0 = f->nd > 1 ? 1 : 0, j1 = f->nd > 1 ? f->d[1] : 0;
    #pragma omp parallel for collapse(2) reduction(max : m) reduction(| : nanflag) schedule(static)
    for (int64_t k = k0; k <= k1; ++k)
        for (int64_t j = j0; j <= j1; ++j)
            for (int64_t i = 1; i <= f->d[0]; ++i) {
                og_real a = FABS(AT(f, i, j, k));
                if (isnan(a)) nanflag = 1;
                if (a > m) m = a;
            }
    return nanflag ? NAN : m;
}

/* ---------------------------------------------------------------- kernels ----------------- */
/* Index space of every op: src/KernelLaunch.jl:40-41,108-109 -> I = J - 1, J in 1..n+2 ; the caller
 * (oracle.py Launcher restatement) passes the region box. */

#define LOOP3(lo, hi)                                                         \
    _Pragma("omp parallel for collapse(2) schedule(static)")                  \
    for (int64_t k = (lo)[2]; k <= (hi)[2]; ++k)                              \
        for (int64_t j = (lo)[1]; j <= (hi)[1]; ++j)                          \
            for (int64_t i = (lo)[0]; i <= (hi)[0]; ++i)

/* examples/diffusion_2d.jl:8-13 :  q.x = -chi * dx(C) ; q.y = -chi * dy(C) */
void og_compute_q(const og_grid* g, og_field* qx, og_field* qy, const og_field* C, og_real chi,
                  const int64_t* lo, const int64_t* hi) {
    LOOP3(lo, hi) {
        AT(qx, i, j, k) = (-chi) * f_d(g, C, 0, i, j, k);
        AT(qy, i, j, k) = (-chi) * f_d(g, C, 1, i, j, k);
    }
}

/* examples/diffusion_2d.jl:15-19 :  C -= dt * divg(q) ; divg = dx(q.x) + dy(q.y) (field_operators.jl:50-55) */
void og_update_C(const og_grid* g, og_field* C, const og_field* qx, const og_field* qy, og_real dt,
                 const int64_t* lo, const int64_t* hi) {
    LOOP3(lo, hi) {
        og_real dv = f_d(g, qx, 0, i, j, k) + f_d(g, qy, 1, i, j, k);
        AT(C, i, j, k) = AT(C, i, j, k) - dt * dv;
    }
}

/* examples/stokes_3d_inc_ve_T.jl:11-21 / stokes_2d_inc_ve_T.jl:11-18 : dst[I] = src[I], same I for every pair */
void og_update_old(const og_grid* g, int npairs, og_field* const* dst, const og_field* const* src,
                   const int64_t* lo, const int64_t* hi) {
    (void)g;
    for (int p = 0; p < npairs; ++p) {
        og_field* D = dst[p]; const og_field* S = src[p];
        LOOP3(lo, hi) { AT(D, i, j, k) = AT(S, i, j, k); }
    }
}

/* examples/stokes_2d_inc_ve_T.jl:20-34.  tau = {xx,yy,xy}, V = {x,y} */
void og_update_stress2(const og_grid* g, og_field* const* tau, og_field* Pr, og_field* divV,
                       const og_field* const* V, const og_field* const* tau_old,
                       og_real eta, og_real eta_ve, og_real G, og_real dt, og_real dtau_Pr, og_real dtau_r,
                       const int64_t* lo, const int64_t* hi) {
    LOOP3(lo, hi) {
        og_real exx = f_d(g, V[0], 0, i, j, k);
        og_real eyy = f_d(g, V[1], 1, i, j, k);
        og_wide exy = 0.5 * (f_d(g, V[0], 1, i, j, k) + f_d(g, V[1], 0, i, j, k));
        og_real dv  = f_d(g, V[0], 0, i, j, k) + f_d(g, V[1], 1, i, j, k);          /* divg(V) */
        AT(divV, i, j, k) = dv;
        AT(Pr, i, j, k) = AT(Pr, i, j, k) - dv * eta_ve * dtau_Pr;
        const og_wide e[3] = {exx - dv / 3.0, eyy - dv / 3.0, exy};
        og_wide r[3];
        for (int c = 0; c < 3; ++c) {
            og_real t = AT(tau[c], i, j, k), to = AT(tau_old[c], i, j, k);
            r[c] = -(t - to) / (G * dt) - t / eta + 2.0 * e[c];
        }
        for (int c = 0; c < 3; ++c)
            AT(tau[c], i, j, k) = AT(tau[c], i, j, k) + r[c] * eta_ve * dtau_r;
    }
}

/* examples/stokes_3d_inc_ve_T.jl:23-46.  tau = {xx,yy,zz,xy,xz,yz}, V = {x,y,z} */
void og_update_stress3(const og_grid* g, og_field* const* tau, og_field* Pr, og_field* divV,
                       const og_field* const* V, const og_field* const* tau_old,
                       og_real eta, og_real eta_ve, og_real G, og_real dt, og_real dtau_Pr, og_real dtau_r,
                       const int64_t* lo, const int64_t* hi) {
    LOOP3(lo, hi) {
        og_real exx = f_d(g, V[0], 0, i, j, k);
        og_real eyy = f_d(g, V[1], 1, i, j, k);
        og_real ezz = f_d(g, V[2], 2, i, j, k);
        og_wide exy = 0.5 * (f_d(g, V[0], 1, i, j, k) + f_d(g, V[1], 0, i, j, k));
        og_wide exz = 0.5 * (f_d(g, V[0], 2, i, j, k) + f_d(g, V[2], 0, i, j, k));
        og_wide eyz = 0.5 * (f_d(g, V[1], 2, i, j, k) + f_d(g, V[2], 1, i, j, k));
        og_real dv  = f_d(g, V[0], 0, i, j, k) + f_d(g, V[1], 1, i, j, k) + f_d(g, V[2], 2, i, j, k);
        AT(divV, i, j, k) = dv;
        AT(Pr, i, j, k) = AT(Pr, i, j, k) - dv * eta_ve * dtau_Pr;
        const og_wide e[6] = {exx - dv / 3.0, eyy - dv / 3.0, ezz - dv / 3.0, exy, exz, eyz};
        og_wide r[6];
        for (int c = 0; c < 6; ++c) {
            og_real t = AT(tau[c], i, j, k), to = AT(tau_old[c], i, j, k);
            r[c] = -(t - to) / (G * dt) - t / eta + 2.0 * e[c];
        }
        for (int c = 0; c < 6; ++c)
            AT(tau[c], i, j, k) = AT(tau[c], i, j, k) + r[c] * eta_ve * dtau_r;
    }
}

static inline og_real rhog_at(const og_grid* g, const og_field* rhog, const og_inclusion* inc,
                             int64_t i, int64_t j, int64_t k) {
    if (inc && inc->active) return incl_value(g, inc, inc->loc, i, j, k);   /* function_field.jl:49-59 */
    return AT(rhog, i, j, k);
}

/* examples/stokes_2d_inc_ve_T.jl:36-43 */
void og_update_velocity2(const og_grid* g, og_field* const* V, og_field* const* rV, const og_field* Pr,
                         const og_field* const* tau, const og_field* rhog, const og_inclusion* inc,
                         og_real eta_ve, og_real nudtau, const int64_t* lo, const int64_t* hi) {
    const og_field *txx = tau[0], *tyy = tau[1], *txy = tau[2];
    LOOP3(lo, hi) {
        og_real rx = -f_d(g, Pr, 0, i, j, k) + f_d(g, txx, 0, i, j, k) + f_d(g, txy, 1, i, j, k);
        og_real ry = -f_d(g, Pr, 1, i, j, k) + f_d(g, tyy, 1, i, j, k) + f_d(g, txy, 0, i, j, k)
                    - rhog_at(g, rhog, inc, i, j, k);
        AT(rV[0], i, j, k) = rx;
        AT(rV[1], i, j, k) = ry;
        AT(V[0], i, j, k) = AT(V[0], i, j, k) + rx * nudtau / eta_ve;
        AT(V[1], i, j, k) = AT(V[1], i, j, k) + ry * nudtau / eta_ve;
    }
}

/* examples/stokes_3d_inc_ve_T.jl:48-57 */
void og_update_velocity3(const og_grid* g, og_field* const* V, og_field* const* rV, const og_field* Pr,
                         const og_field* const* tau, const og_field* rhog, const og_inclusion* inc,
                         og_real eta_ve, og_real nudtau, const int64_t* lo, const int64_t* hi) {
    const og_field *txx = tau[0], *tyy = tau[1], *tzz = tau[2], *txy = tau[3], *txz = tau[4], *tyz = tau[5];
    LOOP3(lo, hi) {
        og_real rx = -f_d(g, Pr, 0, i, j, k) + f_d(g, txx, 0, i, j, k) + f_d(g, txy, 1, i, j, k) + f_d(g, txz, 2, i, j, k);
        og_real ry = -f_d(g, Pr, 1, i, j, k) + f_d(g, tyy, 1, i, j, k) + f_d(g, txy, 0, i, j, k) + f_d(g, tyz, 2, i, j, k);
        og_real rz = -f_d(g, Pr, 2, i, j, k) + f_d(g, tzz, 2, i, j, k) + f_d(g, txz, 0, i, j, k) + f_d(g, tyz, 1, i, j, k)
                    - rhog_at(g, rhog, inc, i, j, k);
        AT(rV[0], i, j, k) = rx;
        AT(rV[1], i, j, k) = ry;
        AT(rV[2], i, j, k) = rz;
        AT(V[0], i, j, k) = AT(V[0], i, j, k) + rx * nudtau / eta_ve;
        AT(V[1], i, j, k) = AT(V[1], i, j, k) + ry * nudtau / eta_ve;
        AT(V[2], i, j, k) = AT(V[2], i, j, k) + rz * nudtau / eta_ve;
    }
}

/* examples/stokes_3d_inc_ve_T.jl:59-71 (2D: stokes_2d_inc_ve_T.jl:45-54)
 *   qT.d = -lambda * d_d(T) + max(V.d,0)*left_d(T) + min(V.d,0)*right_d(T) */
void og_update_thermal_flux(const og_grid* g, og_field* const* qT, const og_field* T, const og_field* const* V,
                            og_real lambda, const int64_t* lo, const int64_t* hi) {
    for (int d = 0; d < g->nd; ++d) {
        og_field* q = qT[d]; const og_field* v = V[d];
        LOOP3(lo, hi) {
            const og_wide vv = AT(v, i, j, k);      /* max(V, 0.0): the Float64 literal promotes both arguments */
            AT(q, i, j, k) = (og_real)((-lambda) * f_d(g, T, d, i, j, k) + jl_maxw(vv, 0.0) * f_left(T, d, i, j, k)
                                       + jl_minw(vv, 0.0) * f_right(T, d, i, j, k));
        }
    }
}

/* examples/stokes_3d_inc_ve_T.jl:73-77 :  T = T_old - dt * divg(qT) */
void og_update_thermal(const og_grid* g, og_field* T, const og_field* T_old, const og_field* const* qT,
                       og_real dt, const int64_t* lo, const int64_t* hi) {
    const int nd = g->nd;
    LOOP3(lo, hi) {
        og_real dv = f_d(g, qT[0], 0, i, j, k) + f_d(g, qT[1], 1, i, j, k);
        if (nd > 2) dv = dv + f_d(g, qT[2], 2, i, j, k);
        AT(T, i, j, k) = AT(T_old, i, j, k) - dt * dv;
    }
}

/* ---------------------------------------------------------------- boundary conditions ----- */

/* One field, one (dim, side).  Face index range: src/BoundaryConditions/batch.jl:159-160,180-184
 * (worksize = remove_dim(dim, nvertices + 2), I = J - 1  ->  transverse I_t in 0..n_t+2).
 * Rules: first_order_boundary_condition.jl:50-84 (halo_index = first/lastindex = 1 | d ; itp_halo_index = 0 | d+1)
 *   Dirichlet, Vertex along dim : f[b]  = v                         b  = 1 | d
 *   Dirichlet, Center along dim : f[h]  = muladd(2, v - f[nb], f[nb])   h = 0 | d+1, nb = 1 | d
 *   Neumann                     : f[h]  = muladd(spacing, -/+q, f[nb])                                   */
/* value: nothing -> 0 | Number | lower-dimensional Field indexed by remove_dim(dim, I)
 * (first_order_boundary_condition.jl:34-40, src/utils.jl:27-34); vf == NULL -> the constant `value` */
void og_bc_apply_field(const og_grid* g, og_field* f, int dim, int side, int kind, og_real value, const og_field* vf) {
    int64_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int t = 0; t < g->nd; ++t) hi[t] = g->n[t] + 2;
    const int64_t d = f->d[dim];
    const int64_t b  = side == 0 ? 1 : d;
    const int64_t h  = side == 0 ? 0 : d + 1;
    const og_real  sp = g->spacing[dim];
    lo[dim] = hi[dim] = 0;
    for (int64_t k = lo[2]; k <= hi[2]; ++k)
        for (int64_t j = lo[1]; j <= hi[1]; ++j)
            for (int64_t i = lo[0]; i <= hi[0]; ++i) {
                int64_t I[3] = {i, j, k}, N[3] = {i, j, k};
                og_real v = value;
                if (vf) {   /* bc.value[remove_dim(dim, I)...] */
                    int64_t R[3] = {0, 0, 0};
                    int t = 0;
                    for (int a = 0; a < g->nd; ++a) if (a != dim) R[t++] = I[a];
                    v = AT(vf, R[0], R[1], R[2]);
                }
                const og_real qs = side == 0 ? -v : v;
                if (kind == OG_DIRICHLET && f->loc[dim] == OG_VERTEX) {
                    I[dim] = b;
                    AT(f, I[0], I[1], I[2]) = v;
                } else if (kind == OG_DIRICHLET) {
                    I[dim] = h; N[dim] = b;
                    og_real nb = AT(f, N[0], N[1], N[2]);
                    AT(f, I[0], I[1], I[2]) = FMA(2.0, v - nb, nb);
                } else {
                    I[dim] = h; N[dim] = b;
                    AT(f, I[0], I[1], I[2]) = FMA(sp, qs, AT(f, N[0], N[1], N[2]));
                }
            }
}

void og_bc_apply(const og_grid* g, og_field* f, int dim, int side, int kind, og_real value) {
    og_bc_apply_field(g, f, dim, side, kind, value, 0);
}

/* ---------------------------------------------------------------- halo slabs -------------- */

/* src/Distributed/communication_views.jl:1-34 with halo_width = 1 (logical indices):
 *   recv : side 1 -> 0          side 2 -> d+1
 *   send : side 1 -> 1+overlap  side 2 -> d-overlap      overlap = 1 (Vertex) | 0 (Center)
 * other dims: the entire padded storage extent (Colon()).                                   */
int64_t og_slab_len(const og_field* f, int dim) {
    int64_t len = 1;
    for (int t = 0; t < f->nd; ++t) if (t != dim) len *= f->sd[t];
    return len;
}

static void slab_copy(og_field* f, int dim, int64_t idx, og_real* buf, int pack) {
    int64_t lo[3], hi[3];
    for (int t = 0; t < 3; ++t) { lo[t] = -f->o[t]; hi[t] = f->sd[t] - 1 - f->o[t]; }
    lo[dim] = hi[dim] = idx;
    size_t p = 0;
    for (int64_t k = lo[2]; k <= hi[2]; ++k)
        for (int64_t j = lo[1]; j <= hi[1]; ++j)
            for (int64_t i = lo[0]; i <= hi[0]; ++i, ++p) {
                if (pack) buf[p] = AT(f, i, j, k); else AT(f, i, j, k) = buf[p];
            }
}

void og_pack_send(const og_field* f, int dim, int side, og_real* buf) {
    int64_t ov = f->loc[dim] == OG_VERTEX ? 1 : 0;
    int64_t idx = side == 0 ? 1 + ov : f->d[dim] - ov;
    slab_copy((og_field*)f, dim, idx, buf, 1);
}

void og_unpack_recv(og_field* f, int dim, int side, const og_real* buf) {
    int64_t idx = side == 0 ? 0 : f->d[dim] + 1;
    slab_copy(f, dim, idx, (og_real*)buf, 0);
}
