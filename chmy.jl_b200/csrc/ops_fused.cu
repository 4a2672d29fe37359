// ops_fused.cu -- device side and launch glue of the fused update_stress! + update_velocity! sweep (design, data flow
// and the phase functions: fused_sv.cuh).  One CTA = 32 lanes x TYB warp-rows; CL CTAs stacked along y form a
// thread-block cluster whose members read each other's boundary rows through distributed shared memory, so only the
// first and last row of a cluster recompute stresses for their neighbours.
#include <cooperative_groups.h>

#include "fused_sv.cuh"

namespace cg = cooperative_groups;

__device__ __forceinline__ void fsv_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void fsv_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// relaxed flavour: a CTA-scope fence makes this thread's shared-memory writes performed in the SM's shared memory (the
// single point of coherence of DSMEM reads), then the arrive carries no memory semantics of its own.  The release
// flavour is a GPU-scope fence that also waits for the global stores of phase A (measured: the `membar` stall reason).
__device__ __forceinline__ void fsv_cluster_arrive_relaxed() {
    asm volatile("fence.acq_rel.cta;\n\tbarrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}

template <int TD, bool FUN, int TYB>
__global__ void __launch_bounds__(FSV_LANES* TYB, (TYB == 4 ? 3 : 2)) k_fused_sv(const FusedP p, const int cl, const int variant) {
    extern __shared__ __align__(16) double xb[];
    const int lane = threadIdx.x, ty = threadIdx.y;
    int cr = 0;
    const double *below = xb, *above = xb;
    int rb = ty, ra = ty;
    if (ty > 0) rb = ty - 1;
    if (ty < TYB - 1) ra = ty + 1;
    if (cl > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cr = (int)cluster.block_rank();
        if (ty == 0 && cr > 0) { below = cluster.map_shared_rank(xb, cr - 1); rb = TYB - 1; }
        if (ty == TYB - 1 && cr < cl - 1) { above = cluster.map_shared_rank(xb, cr + 1); ra = 0; }
    }
    const bool relaxed = (variant & 1) != 0;
    int bx, cyc, bz;
    const bool boundary = tile_decode(p.order, (int)blockIdx.x, bx, cyc, bz);     // grid = (clusters, CTAs per cluster, 1)
    FusedT s;
    fsv_init(s, p, lane, ty, cr * TYB + ty, bx, cyc, bz, FUN);
    if (cl > 1) fsv_cluster_arrive();
    for (int kp = s.k0 - 1; kp <= s.k1; ++kp) {
        d2 sn[FSV_NF];
        fsv_phase_a<TD>(s, p, kp, sn);
        // every thread of the cluster has finished reading the buffer that is about to be overwritten, and the
        // stresses of plane kp-1 that phase B reads have been published
        if (cl > 1) fsv_cluster_wait(); else __syncthreads();
        fsv_phase_b<TD, FUN>(s, p, kp, sn, TYB, xb, below, rb, above, ra);
        if (cl > 1) { if (relaxed) fsv_cluster_arrive_relaxed(); else fsv_cluster_arrive(); }
    }
    if (cl > 1) fsv_cluster_wait();   // no CTA may exit while a neighbour can still read its shared memory
    if (boundary && p.done) {         // tell the boundary stream: this CTA's part of the outer cells is in memory
        __syncthreads();
        if (lane == 0 && ty == 0) { __threadfence(); atomicAdd(p.done, 1u); }
    }
}

// ---------------------------------------------------------------------------------------------- frame copy
// Cells of a ping-pong field that lie outside the op's index range [0, n+1]^N are never written by the sweep; they
// are carried over from the current buffer to the shadow buffer so that the shadow is a complete field afterwards.
struct FramePair {
    const double* src;   // logical (0,0,0) of the current buffer
    double*       dst;   // ... of the shadow buffer
    int sy, sz;
    int d[3];            // logical field dims
};
struct FrameBatch {
    int       n;
    int       nn[3];     // grid cells per dim: inside = [0, nn+1]
    FramePair f[10];
};

__global__ void __launch_bounds__(256) k_frame_copy(const FrameBatch b) {
    const FramePair& f = b.f[blockIdx.z];
    // six slabs in logical indices [lo, hi] (inclusive); together they tile storage minus [0, n+1]^3
    const int slab = blockIdx.y;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = -1; hi[a] = f.d[a] + 2; }
    const int D = 2 - slab / 2, side = slab & 1;               // z slabs first, then y, then x
    for (int a = D + 1; a < 3; ++a) { lo[a] = 0; hi[a] = b.nn[a] + 1; }
    if (side == 0) hi[D] = -1; else lo[D] = b.nn[D] + 2;
    const long long ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
    const long long total = ex * ey * ez;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = lo[0] + (int)(t % ex), j = lo[1] + (int)((t / ex) % ey), k = lo[2] + (int)(t / (ex * ey));
        const long long off = (long long)i + (long long)j * f.sy + (long long)k * f.sz;
        f.dst[off] = f.src[off];
    }
}

// ---------------------------------------------------------------------------------------------- host side
// rows per CTA with an instantiation (profiles/r2_c1_tune_fused_767.log: 2-, 12- and 16-row CTAs and the register-pipelined
// flavour lost on the B200 and were removed; so did cache-policy hints and a TMA-fed operand ring, profiles/README.md)
static bool fsv_tyb_ok(int v) { return v == 4 || v == 6 || v == 8; }

// defaults: measured optimum at 767^3 (profiles/r2_c2_tune_fused_clusters_hints.log); the environment overrides them
void chmy_tuning_defaults(chmy_tuning* t) {
    t->fuse_tyb = 6; t->fuse_cl = 4; t->fuse_cz = 64; t->fuse_var = 1;
    t->f2_cy = 0; t->f2_unroll = 0; t->t3_cz = 64;       // 0: per-sweep optimum (ops_fused2d.cu); 64 planes: profiles/r2_c20_*
    t->overlap = 1; t->bc_fold = 1;
    const char* e;
    if ((e = getenv("CHMY_FUSE_VARIANT"))) t->fuse_var = atoi(e);
    if ((e = getenv("CHMY_FUSE_TYB")) && fsv_tyb_ok(atoi(e))) t->fuse_tyb = atoi(e);
    if ((e = getenv("CHMY_FUSE_CL")) && atoi(e) >= 1 && atoi(e) <= 8) t->fuse_cl = atoi(e);
    if ((e = getenv("CHMY_FUSE_CZ")) && atoi(e) >= 1) t->fuse_cz = atoi(e);
    if ((e = getenv("CHMY_FUSE2D_CY")) && atoi(e) >= 1) t->f2_cy = atoi(e);
    if ((e = getenv("CHMY_FUSE2D_UNROLL")) && (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4)) t->f2_unroll = atoi(e);
    if ((e = getenv("CHMY_FUSE_T3_CZ")) && atoi(e) >= 1) t->t3_cz = atoi(e);
    if ((e = getenv("CHMY_OVERLAP")) && (e[0] == '0' || e[0] == '1')) t->overlap = e[0] - '0';
    if ((e = getenv("CHMY_BC_FOLD")) && (e[0] == '0' || e[0] == '1')) t->bc_fold = e[0] - '0';
}

extern "C" int chmy_set_fused_tuning(chmy_ctx* ctx, int rows_per_cta, int cluster_size, int z_chunk, int variant) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    if (variant >= 0) ctx->tun.fuse_var = variant;
    if (rows_per_cta > 0) {
        CHMY_REQUIRE(fsv_tyb_ok(rows_per_cta), "rows_per_cta must be 4, 6 or 8");
        ctx->tun.fuse_tyb = rows_per_cta;
    }
    if (cluster_size > 0) {
        CHMY_REQUIRE(cluster_size >= 1 && cluster_size <= 8, "cluster_size must be 1..8 (the portable cluster limit)");
        ctx->tun.fuse_cl = cluster_size;
    }
    if (z_chunk > 0) ctx->tun.fuse_cz = z_chunk;
    return CHMY_OK;
}

template <int TD, bool FUN, int TYB>
static int launch_fused(const FusedP& p, int cl, int variant, dim3 grid, cudaStream_t st) {
    void (*kern)(const FusedP, const int, const int) = k_fused_sv<TD, FUN, TYB>;
    const size_t smem = fsv_smem_bytes(TYB);
    static bool attr_done[64] = {};   // per instantiation and device (function attributes are per device)
    int dev = 0;
    CHMY_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CHMY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(FSV_LANES, TYB, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)cl; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = cl > 1 ? 1 : 0;
    CHMY_CUDA(cudaLaunchKernelEx(&cfg, kern, p, cl, variant));
    return CHMY_OK;
}

template <int TD, bool FUN>
static int launch_fused_tyb(const FusedP& p, int tyb, int cl, int variant, dim3 grid, cudaStream_t st) {
    switch (tyb) {
    case 4: return launch_fused<TD, FUN, 4>(p, cl, variant, grid, st);
    case 6: return launch_fused<TD, FUN, 6>(p, cl, variant, grid, st);
    default: return launch_fused<TD, FUN, 8>(p, cl, variant, grid, st);
    }
}

// fields of the two descriptors (ops.cu / include/chmy_b200.h):
//   stress  : tau[6] Pr divV V[3] tau_old[6]   scalars eta eta_ve G dt dtau_Pr dtau_r
//   velocity: V[3] r_V[3] Pr tau[6] rho_g|NULL  scalars eta_ve nudtau
bool chmy_fused_eligible(const chmy_launch_desc* ds, const chmy_launch_desc* dv) {
    if (chmy_fast_disabled()) return false;
    if (ds->grid.ndims != 3 || dv->grid.ndims != 3) return false;
    for (int a = 0; a < 3; ++a)
        if (ds->grid.n[a] != dv->grid.n[a] || ds->grid.inv_spacing[a] != dv->grid.inv_spacing[a]) return false;
    chmy_field* const* S = ds->fields;
    chmy_field* const* V = dv->fields;
    for (int c = 0; c < 6; ++c)
        if (S[c] != V[7 + c]) return false;
    if (S[6] != V[6]) return false;
    for (int c = 0; c < 3; ++c)
        if (S[8 + c] != V[c]) return false;
    if (ds->scalars[1] != dv->scalars[0]) return false;        // eta_ve
    for (int q = 0; q < ds->nfields; ++q)
        if (!S[q] || !aligned16(S[q])) return false;
    for (int q = 0; q < dv->nfields; ++q)
        if (V[q] && !aligned16(V[q])) return false;
    const chmy_field *CC = S[0], *VC = S[4], *CV = S[5], *rho = V[13];
    if (!same_strides(S[1], CC) || !same_strides(S[2], CC) || !same_strides(S[6], CC) || !same_strides(S[7], CC) ||
        !same_strides(S[10], CC) || !same_strides(S[8], VC) || !same_strides(S[9], CV) || !same_strides(V[5], CC) ||
        !same_strides(V[3], VC) || !same_strides(V[4], CV) || (rho && !same_strides(rho, CC)))
        return false;
    for (int c = 0; c < 6; ++c)
        if (!same_strides(S[11 + c], S[c])) return false;
    return true;
}

// One sub-box of the fused op.  cur/shadow pointers are passed explicitly (the caller swaps the fields' buffers).
// tiles of [lo, lo + n) of width w along one dim: [i0, i1) are those that stay `m_lo` cells away from index 0 and `m_hi`
// cells away from index `full` (the op's range is [0, full)); everything else is a boundary tile
static void interior_tiles(int lo, int n, int w, int full, int m_lo, int m_hi, int* g, int* i0, int* i1) {
    *g = (n + w - 1) / w;
    int a = 0, b = *g;
    while (a < *g && lo + a * w < m_lo) ++a;
    while (b > a && (lo + b * w < lo + n ? lo + b * w : lo + n) > full - m_hi) --b;
    *i0 = a; *i1 = b;
}

extern "C" int chmy_selftest_tile_order(const int32_t g[3], const int32_t i0[3], const int32_t i1[3], int32_t tail, int32_t* out /* 4 per tile */) {
    CHMY_REQUIRE(g && i0 && i1 && out, "NULL argument");
    TileOrder o;
    for (int a = 0; a < 3; ++a) { o.g[a] = g[a]; o.i0[a] = i0[a]; o.i1[a] = i1[a]; }
    o.tail = tail ? 1 : 0;
    for (int c = 0; c < tile_total(o); ++c) {
        int bx, by, bz;
        const bool bnd = tile_decode(o, c, bx, by, bz);
        out[4 * c] = bx; out[4 * c + 1] = by; out[4 * c + 2] = bz; out[4 * c + 3] = bnd ? 1 : 0;
    }
    return CHMY_OK;
}

// div2_exact() per divisor, remembered per context (the drivers launch with the same four scalars every iteration)
bool chmy_div2_cached(chmy_ctx* ctx, double c) {
    for (int q = 0; q < ctx->n_div2; ++q)
        if (ctx->div2_c[q] == c) return ctx->div2_ok[q];
    const bool ok = div2_exact(c);
    const int q = ctx->n_div2 < 8 ? ctx->n_div2++ : (ctx->div2_next++ & 7);
    ctx->div2_c[q] = c; ctx->div2_ok[q] = ok;
    return ok;
}

extern "C" int chmy_division_two_op_exact(double c, int32_t* exact) {
    CHMY_REQUIRE(exact != nullptr, "NULL argument");
    *exact = div2_exact(c) ? 1 : 0;
    return CHMY_OK;
}

extern "C" int chmy_last_division_mode(const chmy_ctx* ctx, int32_t* mode) {
    CHMY_REQUIRE(ctx && mode, "NULL argument");
    *mode = ctx->div_mode;
    return CHMY_OK;
}

int chmy_run_fused(chmy_ctx* ctx, const chmy_launch_desc* ds, const chmy_launch_desc* dv, const Box& box,
                   double* const* cur /* tau[6] Pr V[3] */, double* const* shadow, cudaStream_t st, unsigned int* done,
                   unsigned int* n_signal) {
    if (n_signal) *n_signal = 0;
    if (box.n[0] <= 0 || box.n[1] <= 0 || box.n[2] <= 0) return CHMY_OK;
    CHMY_REQUIRE((box.lo[0] & 1) == 0, "fused sweep needs an even x origin");
    chmy_field* const* S = ds->fields;
    chmy_field* const* V = dv->fields;
    const double* s = ds->scalars;
    const double* id = ds->grid.inv_spacing;
    FusedP p;
    memset(&p, 0, sizeof(p));
    for (int c = 0; c < 6; ++c) { p.tc[c] = cur[c]; p.tn[c] = shadow[c]; p.to[c] = S[11 + c]->p0; }
    p.Prc = cur[6]; p.Prn = shadow[6];
    for (int c = 0; c < 3; ++c) { p.Vc[c] = cur[7 + c]; p.Vn[c] = shadow[7 + c]; p.r[c] = V[3 + c]->p0; }
    p.dV = S[7]->p0;
    const chmy_field* rho = V[13];
    p.rho = rho ? rho->p0 : nullptr;
    p.cc = strides_of(S[0]); p.vv = strides_of(S[3]); p.vc = strides_of(S[4]); p.cv = strides_of(S[5]);
    for (int a = 0; a < 3; ++a) {
        p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a];
        p.flo[a] = 0; p.fhi[a] = (int)ds->grid.n[a] + 2;
    }
    p.idx = id[0]; p.idy = id[1]; p.idz = id[2];
    p.eta_ve = s[1]; p.dtau_Pr = s[4]; p.dtau_r = s[5]; p.nudtau = dv->scalars[1];
    const double Gdt = s[2] * s[3];
    p.Gdt = divc_of(Gdt); p.eta = divc_of(s[0]); p.three = divc_of(3.0);
    p.eve = divc_of(s[1]);
    if (!rho) {
        p.inc.active = 1; p.inc.nd = 3;
        for (int a = 0; a < 3; ++a) {
            p.inc.loc[a] = dv->rho_g.loc[a]; p.inc.origin[a] = dv->grid.origin[a];
            p.inc.spacing[a] = dv->grid.spacing[a]; p.inc.c0[a] = dv->rho_g.c0[a];
        }
        p.inc.r2 = dv->rho_g.r * dv->rho_g.r; p.inc.in = dv->rho_g.in; p.inc.out = dv->rho_g.out;
    }
    // division mode (fast_common.cuh): true division when a divisor is outside Markstein's conditions (or on request), the
    // two-operation sequence when it is PROVEN exact for all four divisors of this launch, else the four-operation sequence
    const bool td = chmy_force_true_div() || !markstein_ok(Gdt) || !markstein_ok(s[0]) || !markstein_ok(s[1]);
    static const bool allow2 = !(getenv("CHMY_DIV2") && getenv("CHMY_DIV2")[0] == '0');
    const int dm = td ? 1 : (allow2 && chmy_div2_cached(ctx, Gdt) && chmy_div2_cached(ctx, s[0]) && chmy_div2_cached(ctx, s[1]) && chmy_div2_cached(ctx, 3.0)) ? 2 : 0;
    ctx->div_mode = dm;
    // geometry: clusters shrink for short boxes (slabs of a split launch)
    int tyb = ctx->tun.fuse_tyb, cl = ctx->tun.fuse_cl;
    while (cl > 1 && (cl - 1) * tyb - 2 >= box.n[1]) cl -= 1;
    if (tyb > 4 && cl == 1 && 4 - 2 >= box.n[1]) tyb = 4;
    p.rows_int = cl * tyb - 2;
    const int nch = (box.n[2] + ctx->tun.fuse_cz - 1) / ctx->tun.fuse_cz;
    p.cz = (box.n[2] + nch - 1) / nch;
    // boundary tiles: whatever owns a cell the batches / the exchange read or write -- indices 0..2 and n-2..n+1 of a dim
    // (halo, boundary node, first/last interior cell, send planes).  Without a counter: natural order, nobody signals.
    const int tw[3] = {FSV_XI, p.rows_int, p.cz};
    for (int a = 0; a < 3; ++a) {
        if (done) interior_tiles(box.lo[a], box.n[a], tw[a], p.fhi[a], 3, 4, &p.order.g[a], &p.order.i0[a], &p.order.i1[a]);
        else { p.order.g[a] = (box.n[a] + tw[a] - 1) / tw[a]; p.order.i0[a] = 0; p.order.i1[a] = p.order.g[a]; }
    }
    p.order.tail = done ? 1 : 0;
    p.done = done;
    const long long total = (long long)p.order.g[0] * p.order.g[1] * p.order.g[2];
    CHMY_REQUIRE(total < (1ll << 31), "too many tiles for one sweep");
    if (n_signal) *n_signal = (unsigned int)((tile_total(p.order) - tile_interior(p.order)) * cl);
    const dim3 grid((unsigned)total, (unsigned)cl, 1);
    int rc;
    const int var = ctx->tun.fuse_var;
    if (ctx->sweep_ev0) CHMY_TRY(chmy_event_record_on(ctx, ctx->sweep_ev0 - 1, st));
    if (rho) rc = dm == 1 ? launch_fused_tyb<1, false>(p, tyb, cl, var, grid, st)
                : dm == 2 ? launch_fused_tyb<2, false>(p, tyb, cl, var, grid, st) : launch_fused_tyb<0, false>(p, tyb, cl, var, grid, st);
    else     rc = dm == 1 ? launch_fused_tyb<1, true>(p, tyb, cl, var, grid, st)
                : dm == 2 ? launch_fused_tyb<2, true>(p, tyb, cl, var, grid, st) : launch_fused_tyb<0, true>(p, tyb, cl, var, grid, st);
    CHMY_TRY(rc);
    if (ctx->sweep_ev1) CHMY_TRY(chmy_event_record_on(ctx, ctx->sweep_ev1 - 1, st));
    ctx->n_launches++;
    return CHMY_OK;
}

// carries the cells outside [0, n+1]^3 of the listed fields from src to dst (one launch)
int chmy_frame_copy(chmy_ctx* ctx, const chmy_grid_desc* g, int n, chmy_field* const* fs, double* const* src, double* const* dst,
                    cudaStream_t st) {
    if (n <= 0) return CHMY_OK;
    CHMY_REQUIRE(n <= 10, "too many fields for one frame copy");
    FrameBatch b;
    memset(&b, 0, sizeof(b));
    b.n = n;
    for (int a = 0; a < 3; ++a) b.nn[a] = (int)g->n[a];
    for (int q = 0; q < n; ++q) {
        b.f[q].src = src[q]; b.f[q].dst = dst[q];
        b.f[q].sy = (int)fs[q]->stride[1]; b.f[q].sz = (int)fs[q]->stride[2];
        for (int a = 0; a < 3; ++a) b.f[q].d[a] = (int)fs[q]->d[a];
    }
    k_frame_copy<<<dim3(64, 6, (unsigned)n), 256, 0, st>>>(b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}
