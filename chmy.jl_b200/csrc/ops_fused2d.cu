// ops_fused2d.cu -- EXPERIMENTAL (opt-in: chmy_set_fusion(ctx, 3); proven bit-exact by the host emulation in
// tests/emul/fused_emul2d.cpp, NOT YET RUN ON A GPU): device side and launch glue of the 2D fused sweeps
//   kind 1  update_stress! + update_velocity!        (fused_sv2d.cuh;   24 -> 18 array passes)
//   kind 2  compute_q! + update_C!                   (fused_pairs2d.cuh; 7 ->  4)
//   kind 3  update_thermal_flux! + update_thermal!   (fused_pairs2d.cuh; 9 ->  7)
// and of the 3D thermal pair
//   kind 4  update_thermal_flux! + update_thermal! 3D (fused_thermal3.cuh; 12 -> 9; emulation: tests/emul/fused_emul_t3.cpp)
// Design and data flow: the two headers.  One warp = one 64-cell row segment (60 interior cells) marching along y over
// one y-chunk; warps are independent (no shared memory, no barrier), a CTA is just F2_WARPS consecutive segments.
#include "fused_sv2d.cuh"
#include "fused_pairs2d.cuh"
#include "fused_thermal3.cuh"

constexpr int F2_WARPS = 4;
constexpr int FT3_ROWS = 8;      // rows (warps) per CTA of the 3D thermal sweep, as the tuned kernels' TY

template <int TD, bool FUN>      // TD: division mode (fast_common.cuh div_m: 0 four operations, 1 div.rn.f64, 2 two operations)
__global__ void __launch_bounds__(FSV_LANES* F2_WARPS, 4) k_fused_sv2(const Fused2P p, const int gx) {
    const int lane = threadIdx.x;
    const int seg  = blockIdx.x * F2_WARPS + threadIdx.y;
    if (seg >= gx) return;                    // whole warps leave; the kernel has no CTA-wide barrier
    Fused2T s;
    fsv2_init(s, p, lane, seg, blockIdx.y, FUN);
    for (int jp = s.j0 - 1; jp <= s.j1; ++jp) {
        d2 sn[4];
        fsv2_phase_a<TD>(s, p, jp, sn);
        // x-neighbours of the NEW values of row jp-1 (the carried registers, read before phase B rotates them)
        const double pr_im1  = __shfl_up_sync(FULL, s.prC.y, 1);
        const double txx_im1 = __shfl_up_sync(FULL, s.txxC.y, 1);
        const double txy_ip2 = __shfl_down_sync(FULL, s.txyC.x, 1);
        fsv2_phase_b<TD, FUN>(s, p, jp, sn, pr_im1, txx_im1, txy_ip2);
    }
}

// U rows per group: the operands of the whole group are requested before its first row is computed (U loads per array
// in flight per thread; the stores of a row cannot be reordered behind later loads by the compiler, they may alias)
template <int KIND, int U>
__global__ void __launch_bounds__(FSV_LANES* F2_WARPS, (KIND == 0 ? 8 : 4)) k_fused_q2(const FusedQ2P p, const int gx) {
    const int lane = threadIdx.x;
    const int seg  = blockIdx.x * F2_WARPS + threadIdx.y;
    if (seg >= gx) return;
    FusedQ2T s;
    fq2_init(s, p, lane, seg, blockIdx.y);
    for (int jg = s.j0; jg <= s.j1; jg += U) {
        FusedQ2L L[U];
#pragma unroll
        for (int r = 0; r < U; ++r)
            if (jg + r <= s.j1) fq2_load<KIND>(s, p, r, L[r]);
#pragma unroll
        for (int r = 0; r < U; ++r) {
            if (jg + r <= s.j1) {             // warp-uniform
                d2 sn[2];
                fq2_phase_a<KIND>(s, p, jg + r, L[r], sn);
                const double qx_ip2 = __shfl_down_sync(FULL, s.qxC.x, 1);
                fq2_phase_b<KIND>(s, p, jg + r, sn, qx_ip2);
            }
        }
    }
}

// 3D thermal pair: one warp = one 64-cell row segment of one row, FT3_ROWS rows per CTA, one z-chunk per CTA; warps are
// independent (no shared memory, no barrier).  All 32 lanes run the loop: its bounds depend on blockIdx.z only.
// Three resident CTAs per SM: 80 registers and 24 warps instead of 112 and 16 (20 bytes of spills, L1 hits -- this sweep has
// no cluster barrier that would empty L1): 6.5 -> 5.7 ms at 767^3; four (64 registers, 124 bytes of spills) gives it back
// (profiles/r2_c45_thermal3_occupancy.log).  The 2D stress+velocity sweep is the other way round: 4 CTAs of 122 registers beat
// 5 or 6 CTAs with spills (r2_c46_stokes2d_sweep_occupancy.log).
__global__ void __launch_bounds__(FSV_LANES* FT3_ROWS, 3) k_fused_t3(const FusedT3P p) {
    FusedT3T s;
    ft3_init(s, p, threadIdx.x, blockIdx.x, blockIdx.y * FT3_ROWS + threadIdx.y, blockIdx.z);
    for (int k = s.k0; k < s.k1; ++k) {
        FusedT3L L;
        ft3_load(s, p, k, L);
        const double t_left   = __shfl_up_sync(FULL, s.t_k.y, 1);
        const double t_right  = __shfl_down_sync(FULL, s.t_k.x, 1);
        const double vx_right = __shfl_down_sync(FULL, L.vx.x, 1);
        ft3_compute(s, p, k, L, t_left, t_right, vx_right);
    }
}

// ---------------------------------------------------------------------------------------------- frame copy (2D)
// Cells of a ping-pong field outside the ops' index range [0, n+1]^2 are never written by a sweep; they are carried
// over from the current buffer to the shadow buffer so that the shadow is a complete field afterwards.
struct Frame2Pair {
    const double* src;   // logical (0,0) of the current buffer
    double*       dst;   // ... of the shadow buffer
    int sy;
    int d[2];            // logical field dims
};
struct Frame2Batch {
    int        nn[2];    // grid cells per dim: inside = [0, nn+1]
    Frame2Pair f[6];
};

__global__ void __launch_bounds__(256) k_frame_copy2(const Frame2Batch b) {
    const Frame2Pair& f = b.f[blockIdx.z];
    // four slabs in logical indices [lo, hi] (inclusive): y slabs over the whole x extent, then x slabs over rows 0..n+1
    const int slab = blockIdx.y;
    int lo[2] = {-1, -1}, hi[2] = {f.d[0] + 2, f.d[1] + 2};
    const int D = 1 - slab / 2, side = slab & 1;
    if (D == 0) { lo[1] = 0; hi[1] = b.nn[1] + 1; }
    if (side == 0) hi[D] = -1; else lo[D] = b.nn[D] + 2;
    const long long ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1;
    const long long total = ex * ey;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = lo[0] + (int)(t % ex), j = lo[1] + (int)(t / ex);
        const long long off = (long long)i + (long long)j * f.sy;
        f.dst[off] = f.src[off];
    }
}

int chmy_frame_copy2(chmy_ctx* ctx, const chmy_grid_desc* g, int n, chmy_field* const* fs, double* const* src, double* const* dst,
                     cudaStream_t st) {
    if (n <= 0) return CHMY_OK;
    CHMY_REQUIRE(n <= 6, "too many fields for one 2D frame copy");
    Frame2Batch b;
    memset(&b, 0, sizeof(b));
    for (int a = 0; a < 2; ++a) b.nn[a] = (int)g->n[a];
    for (int q = 0; q < n; ++q) {
        b.f[q].src = src[q]; b.f[q].dst = dst[q];
        b.f[q].sy = (int)fs[q]->stride[1];
        for (int a = 0; a < 2; ++a) b.f[q].d[a] = (int)fs[q]->d[a];
    }
    k_frame_copy2<<<dim3(32, 4, (unsigned)n), 256, 0, st>>>(b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- host side
extern "C" int chmy_set_fused2d_tuning(chmy_ctx* ctx, int rows_per_chunk, int unroll, int thermal3_planes_per_chunk) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    if (thermal3_planes_per_chunk > 0) ctx->tun.t3_cz = thermal3_planes_per_chunk;
    if (rows_per_chunk > 0) ctx->tun.f2_cy = rows_per_chunk;      // (env / default 0: the per-sweep optimum)
    if (unroll > 0) {
        CHMY_REQUIRE(unroll == 1 || unroll == 2 || unroll == 4, "unroll must be 1, 2 or 4");
        ctx->tun.f2_unroll = unroll;
    }
    return CHMY_OK;
}

static inline bool fits_int(const chmy_field* f) { return f->stride[1] * f->sd[1] < (1ll << 31); }

// Which fused 2D sweep can run the deferred launch `dp` together with `dc`?  0 = none.  Descriptor layouts (ops.cu):
//   stress  : tau[3] Pr divV V[2] tau_old[3]  scalars eta eta_ve G dt dtau_Pr dtau_r ; velocity: V[2] r_V[2] Pr tau[3] rho_g|NULL
//   compute_q: q.x q.y C  scalars chi        ; update_C: C q.x q.y  scalars dt
//   thermal_flux: qT[2] T V[2]  scalars lambda ; thermal: T T_old qT[2]  scalars dt
static inline bool fits_int3(const chmy_field* f) { return f->stride[2] * f->sd[2] < (1ll << 40) && f->stride[2] < (1ll << 31); }

// 3D thermal pair (kind 4).  thermal_flux: qT[3] T V[3]  scalars lambda ; thermal: T T_old qT[3]  scalars dt
static int fused_t3_kind(const chmy_launch_desc* dp, const chmy_launch_desc* dc) {
    if (dp->op != CHMY_OP_UPDATE_THERMAL_FLUX || dc->op != CHMY_OP_UPDATE_THERMAL) return 0;
    chmy_field* const* P = dp->fields;
    chmy_field* const* Q = dc->fields;
    for (int a = 0; a < 3; ++a)
        if (dp->grid.n[a] != dc->grid.n[a] || dp->grid.inv_spacing[a] != dc->grid.inv_spacing[a]) return 0;
    for (int q = 0; q < 7; ++q)
        if (!P[q] || !aligned16(P[q]) || !fits_int3(P[q])) return 0;
    for (int q = 0; q < 5; ++q)
        if (!Q[q] || !aligned16(Q[q]) || !fits_int3(Q[q])) return 0;
    if (P[0] != Q[2] || P[1] != Q[3] || P[2] != Q[4] || P[3] != Q[0]) return 0;       // same qT, same T
    if (Q[1] == Q[0]) return 0;                                                         // T_old must not alias T
    const chmy_field *T = Q[0];
    // storage classes: CC T T_old V.z q.z ; VC V.x q.x ; CV V.y q.y
    if (!same_strides(Q[1], T) || !same_strides(P[2], T) || !same_strides(P[6], T) || !same_strides(P[4], P[0]) ||
        !same_strides(P[5], P[1]))
        return 0;
    return 4;
}

int chmy_fused2d_kind(const chmy_launch_desc* dp, const chmy_launch_desc* dc) {
    if (chmy_fast_disabled()) return 0;
    if (dp->has_bc) return 0;
    if (dp->grid.ndims == 3 && dc->grid.ndims == 3) return fused_t3_kind(dp, dc);
    if (dp->grid.ndims != 2 || dc->grid.ndims != 2) return 0;
    for (int a = 0; a < 2; ++a)
        if (dp->grid.n[a] != dc->grid.n[a] || dp->grid.inv_spacing[a] != dc->grid.inv_spacing[a]) return 0;
    chmy_field* const* P = dp->fields;
    chmy_field* const* Q = dc->fields;
    for (int q = 0; q < dp->nfields; ++q)
        if (!P[q] || !aligned16(P[q]) || !fits_int(P[q])) return 0;
    for (int q = 0; q < dc->nfields; ++q)
        if (Q[q] && (!aligned16(Q[q]) || !fits_int(Q[q]))) return 0;
    if (dp->op == CHMY_OP_UPDATE_STRESS && dc->op == CHMY_OP_UPDATE_VELOCITY) {
        for (int c = 0; c < 3; ++c)
            if (P[c] != Q[5 + c]) return 0;
        if (P[3] != Q[4] || P[5] != Q[0] || P[6] != Q[1]) return 0;
        if (dp->scalars[1] != dc->scalars[0]) return 0;        // eta_ve
        const chmy_field *CC = P[0], *VC = P[5], *CV = P[6], *rho = Q[8];
        if (!same_strides(P[1], CC) || !same_strides(P[3], CC) || !same_strides(P[4], CC) || !same_strides(Q[2], VC) ||
            !same_strides(Q[3], CV) || (rho && !same_strides(rho, CV)))
            return 0;
        for (int c = 0; c < 3; ++c)
            if (!same_strides(P[7 + c], P[c])) return 0;
        return 1;
    }
    if (dp->op == CHMY_OP_COMPUTE_Q && dc->op == CHMY_OP_UPDATE_C) {
        if (P[0] != Q[1] || P[1] != Q[2] || P[2] != Q[0]) return 0;
        return 2;
    }
    if (dp->op == CHMY_OP_UPDATE_THERMAL_FLUX && dc->op == CHMY_OP_UPDATE_THERMAL) {
        if (P[0] != Q[2] || P[1] != Q[3] || P[2] != Q[0]) return 0;
        if (!same_strides(Q[1], Q[0]) || !same_strides(P[3], P[0]) || !same_strides(P[4], P[1])) return 0;
        if (Q[1] == Q[0]) return 0;                             // T_old must not alias T
        return 3;
    }
    return 0;
}

// the fields a sweep of `kind` writes through shadow buffers: tau[3] Pr V[2] | C | T
int chmy_fused2d_pingpong(int kind, const chmy_launch_desc* dp, const chmy_launch_desc* dc, chmy_field** pp) {
    if (kind == 1) {
        for (int c = 0; c < 3; ++c) pp[c] = dp->fields[c];
        pp[3] = dp->fields[3];
        pp[4] = dc->fields[0]; pp[5] = dc->fields[1];
        return 6;
    }
    pp[0] = dc->fields[0];
    return 1;
}

template <int KIND>
static int launch_q2(const FusedQ2P& p, int gx, dim3 grid, int unroll, cudaStream_t st) {
    switch (unroll) {
    case 1: k_fused_q2<KIND, 1><<<grid, dim3(FSV_LANES, F2_WARPS, 1), 0, st>>>(p, gx); break;
    case 2: k_fused_q2<KIND, 2><<<grid, dim3(FSV_LANES, F2_WARPS, 1), 0, st>>>(p, gx); break;
    default: k_fused_q2<KIND, 4><<<grid, dim3(FSV_LANES, F2_WARPS, 1), 0, st>>>(p, gx); break;
    }
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

// One sub-box of a fused 2D sweep.  cur / shadow: buffers of the ping-pong fields in chmy_fused2d_pingpong order (the
// caller has already swapped the fields' storage).

static int run_fused_t3(chmy_ctx* ctx, const chmy_launch_desc* dp, const chmy_launch_desc* dc, const Box& box,
                        double* const* cur, double* const* shadow, cudaStream_t st) {
    if (box.n[0] <= 0 || box.n[1] <= 0 || box.n[2] <= 0) return CHMY_OK;
    CHMY_REQUIRE((box.lo[0] & 1) == 0, "fused sweep needs an even x origin");
    chmy_field* const* P = dp->fields;
    chmy_field* const* Q = dc->fields;
    FusedT3P p;
    memset(&p, 0, sizeof(p));
    p.Tc = cur[0]; p.Tn = shadow[0]; p.To = Q[1]->p0;
    p.qx = P[0]->p0; p.qy = P[1]->p0; p.qz = P[2]->p0;
    p.Vx = P[4]->p0; p.Vy = P[5]->p0; p.Vz = P[6]->p0;
    p.cc = strides_of(Q[0]); p.vc = strides_of(P[0]); p.cv = strides_of(P[1]);
    for (int a = 0; a < 3; ++a) {
        p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a];
        p.flo[a] = 0; p.fhi[a] = (int)dp->grid.n[a] + 2;
    }
    p.lam = dp->scalars[0]; p.dt = dc->scalars[0];
    p.idx = dp->grid.inv_spacing[0]; p.idy = dp->grid.inv_spacing[1]; p.idz = dp->grid.inv_spacing[2];
    int cz = ctx->tun.t3_cz;      // planes per z-chunk (default: the tuned kernels' measured optimum)
    while ((box.n[2] + cz - 1) / cz > 65535) cz *= 2;
    const int nch = (box.n[2] + cz - 1) / cz;
    p.cz = (box.n[2] + nch - 1) / nch;                          // balanced chunks
    const dim3 grid((unsigned)((box.n[0] + 2 * FSV_LANES - 1) / (2 * FSV_LANES)), (unsigned)((box.n[1] + FT3_ROWS - 1) / FT3_ROWS),
                    (unsigned)((box.n[2] + p.cz - 1) / p.cz));
    k_fused_t3<<<grid, dim3(FSV_LANES, FT3_ROWS, 1), 0, st>>>(p);
    CHMY_CUDA(cudaGetLastError());
    ctx->n_launches++;
    return CHMY_OK;
}

int chmy_run_fused2d(chmy_ctx* ctx, int kind, const chmy_launch_desc* dp, const chmy_launch_desc* dc, const Box& box,
                     double* const* cur, double* const* shadow, cudaStream_t st) {
    if (kind == 4) return run_fused_t3(ctx, dp, dc, box, cur, shadow, st);
    if (box.n[0] <= 0 || box.n[1] <= 0) return CHMY_OK;
    CHMY_REQUIRE((box.lo[0] & 1) == 0, "fused sweep needs an even x origin");
    chmy_field* const* P = dp->fields;
    chmy_field* const* Q = dc->fields;
    const double* id = dp->grid.inv_spacing;
    const int gx = (box.n[0] + FSV_XI - 1) / FSV_XI;
    // 0 = the measured optimum of the sweep (profiles/r2_c20_tune_pairs_*.log, 8191^2 / 16383^2): compute_q! + update_C!
    // 16 rows per chunk with 4 rows of loads in flight; 2D Stokes and the 2D thermal pair 32 rows, 1 row in flight
    int cy = ctx->tun.f2_cy > 0 ? ctx->tun.f2_cy : (kind == 2 ? 16 : 32);
    const int unroll = ctx->tun.f2_unroll > 0 ? ctx->tun.f2_unroll : (kind == 2 ? 4 : 1);
    while ((box.n[1] + cy - 1) / cy > 65535) cy *= 2;
    const int nch = (box.n[1] + cy - 1) / cy;
    cy = (box.n[1] + nch - 1) / nch;                            // balanced chunks
    const dim3 grid((unsigned)((gx + F2_WARPS - 1) / F2_WARPS), (unsigned)((box.n[1] + cy - 1) / cy), 1);
    if (kind == 1) {
        const double* s = dp->scalars;
        Fused2P p;
        memset(&p, 0, sizeof(p));
        for (int c = 0; c < 3; ++c) { p.tc[c] = cur[c]; p.tn[c] = shadow[c]; p.to[c] = P[7 + c]->p0; }
        p.Prc = cur[3]; p.Prn = shadow[3];
        for (int c = 0; c < 2; ++c) { p.Vc[c] = cur[4 + c]; p.Vn[c] = shadow[4 + c]; p.r[c] = Q[2 + c]->p0; }
        p.dV = P[4]->p0;
        const chmy_field* rho = Q[8];
        p.rho = rho ? rho->p0 : nullptr;
        p.s_cc = (int)P[0]->stride[1]; p.s_vv = (int)P[2]->stride[1]; p.s_vc = (int)P[5]->stride[1]; p.s_cv = (int)P[6]->stride[1];
        for (int a = 0; a < 2; ++a) {
            p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a];
            p.flo[a] = 0; p.fhi[a] = (int)dp->grid.n[a] + 2;
        }
        p.idx = id[0]; p.idy = id[1];
        p.eta_ve = s[1]; p.dtau_Pr = s[4]; p.dtau_r = s[5]; p.nudtau = dc->scalars[1];
        const double Gdt = s[2] * s[3];
        p.Gdt = divc_of(Gdt); p.eta = divc_of(s[0]); p.three = divc_of(3.0);
        p.eve = divc_of(s[1]);
        if (!rho) {
            p.inc.active = 1; p.inc.nd = 2;
            for (int a = 0; a < 2; ++a) {
                p.inc.loc[a] = dc->rho_g.loc[a]; p.inc.origin[a] = dc->grid.origin[a];
                p.inc.spacing[a] = dc->grid.spacing[a]; p.inc.c0[a] = dc->rho_g.c0[a];
            }
            p.inc.r2 = dc->rho_g.r * dc->rho_g.r; p.inc.in = dc->rho_g.in; p.inc.out = dc->rho_g.out;
        }
        p.cy = cy;
        const bool td = chmy_force_true_div() || !markstein_ok(Gdt) || !markstein_ok(s[0]) || !markstein_ok(s[1]);
        const dim3 blk(FSV_LANES, F2_WARPS, 1);
        static const bool allow2 = !(getenv("CHMY_DIV2") && getenv("CHMY_DIV2")[0] == '0');
        const int dm = td ? 1 : (allow2 && chmy_div2_cached(ctx, Gdt) && chmy_div2_cached(ctx, s[0]) && chmy_div2_cached(ctx, s[1]) && chmy_div2_cached(ctx, 3.0)) ? 2 : 0;
        ctx->div_mode = dm;
        if (rho) {
            if (dm == 1) k_fused_sv2<1, false><<<grid, blk, 0, st>>>(p, gx);
            else if (dm == 2) k_fused_sv2<2, false><<<grid, blk, 0, st>>>(p, gx);
            else k_fused_sv2<0, false><<<grid, blk, 0, st>>>(p, gx);
        } else {
            if (dm == 1) k_fused_sv2<1, true><<<grid, blk, 0, st>>>(p, gx);
            else if (dm == 2) k_fused_sv2<2, true><<<grid, blk, 0, st>>>(p, gx);
            else k_fused_sv2<0, true><<<grid, blk, 0, st>>>(p, gx);
        }
        CHMY_CUDA(cudaGetLastError());
    } else {
        FusedQ2P p;
        memset(&p, 0, sizeof(p));
        p.Cc = cur[0]; p.Cn = shadow[0];
        if (kind == 2) {
            p.qx = P[0]->p0; p.qy = P[1]->p0;
            p.coef = dp->scalars[0]; p.dt = dc->scalars[0];
        } else {
            p.qx = P[0]->p0; p.qy = P[1]->p0; p.Vx = P[3]->p0; p.Vy = P[4]->p0; p.base = Q[1]->p0;
            p.coef = dp->scalars[0]; p.dt = dc->scalars[0];
        }
        p.s_cc = (int)Q[0]->stride[1]; p.s_vc = (int)P[0]->stride[1]; p.s_cv = (int)P[1]->stride[1];
        for (int a = 0; a < 2; ++a) {
            p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a];
            p.flo[a] = 0; p.fhi[a] = (int)dp->grid.n[a] + 2;
        }
        p.idx = id[0]; p.idy = id[1];
        p.cy = cy;
        if (kind == 2) CHMY_TRY(launch_q2<0>(p, gx, grid, unroll, st)); else CHMY_TRY(launch_q2<1>(p, gx, grid, unroll, st));
    }
    ctx->n_launches++;
    return CHMY_OK;
}
