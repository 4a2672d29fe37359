// ops_fused.cu -- device side and launch glue of the fused update_stress! + update_velocity! sweep (design, data flow
// and the phase functions: fused_sv.cuh).  One CTA = 32 lanes x TYB warp-rows; CL CTAs stacked along y form a
// thread-block cluster whose members read each other's boundary rows through distributed shared memory, so only the
// first and last row of a cluster recompute stresses for their neighbours.
#include <cooperative_groups.h>
#include <cuda.h>        // CUtensorMap (types only: the one driver entry point used is resolved at run time)

#include "fused_sv.cuh"
#include "fused_tma.cuh"

namespace cg = cooperative_groups;

__device__ __forceinline__ void fsv_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void fsv_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// relaxed flavour: a CTA-scope fence makes this thread's shared-memory writes performed in the SM's shared memory (the
// single point of coherence of DSMEM reads), then the arrive carries no memory semantics of its own.  The release
// flavour is a GPU-scope fence that also waits for the global stores of phase A (measured: the `membar` stall reason).
__device__ __forceinline__ void fsv_cluster_arrive_relaxed() {
    asm volatile("fence.acq_rel.cta;\n\tbarrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}

template <bool TD, bool FUN, int TYB, int HINT>
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
    FusedT s;
    fsv_init(s, p, lane, ty, cr * TYB + ty, blockIdx.x, blockIdx.y / cl, blockIdx.z, FUN, HINT);
    if (cl > 1) fsv_cluster_arrive();
    for (int kp = s.k0 - 1; kp <= s.k1; ++kp) {
        d2 sn[FSV_NF];
        fsv_phase_a<TD, HINT>(s, p, kp, sn);
        // every thread of the cluster has finished reading the buffer that is about to be overwritten, and the
        // stresses of plane kp-1 that phase B reads have been published
        if (cl > 1) fsv_cluster_wait(); else __syncthreads();
        fsv_phase_b<TD, FUN, HINT>(s, p, kp, sn, TYB, xb, below, rb, above, ra);
        if (cl > 1) { if (relaxed) fsv_cluster_arrive_relaxed(); else fsv_cluster_arrive(); }
    }
    if (cl > 1) fsv_cluster_wait();   // no CTA may exit while a neighbour can still read its shared memory
}

// ---------------------------------------------------------------------------------------------- TMA-fed sweep
// (design: fused_tma.cuh)  mbarrier / bulk-copy primitives, raw PTX for sm_100a
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA unit, SASS UBLKCP); completion is reported to the mbarrier as `bytes` transactions
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// tensor-map flavour of the same row copy: a 64 x 1 x 1 box of a rank-3 tensor map (SASS UTMALDG); out-of-range cells are
// zero-filled by the TMA unit, and all eleven operands share ONE coordinate triple -- a handful of instructions per row
struct FtmMaps {
    CUtensorMap m[FTM_NS];
};
__device__ __forceinline__ void tma_row_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}

// the elected lane of warp 0 requests the CTA's TYB rows of plane z (tensor coordinate) of all eleven operands into slot `slot`
template <int TYB>
__device__ __forceinline__ void ftm_issue(const FtmMaps& maps, double* ring, uint64_t* full, int cx, int cy, int z, int slot) {
    mbar_expect_tx(full, (uint32_t)(FTM_NS * TYB * 64 * sizeof(double)));
#pragma unroll
    for (int op = 0; op < FTM_NS; ++op) tma_row_3d(ring + ftm_ring_off(TYB, slot, op, 0), &maps.m[op], cx, cy, z, full);
}

template <bool TD, bool FUN, int TYB>
__global__ void __launch_bounds__(FSV_LANES* TYB, (TYB == 4 ? 3 : 2)) k_fused_tma(const FusedP p, const int cl, const __grid_constant__ FtmMaps maps) {
    extern __shared__ __align__(128) double sm[];
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int w = __shfl_sync(0xffffffffu, ty, 0);                 // this warp's row, provably warp-uniform
    double*   ring = sm;
    double*   xb   = sm + ftm_xch_base(TYB);
    uint64_t* full  = reinterpret_cast<uint64_t*>(xb + 2 * FSV_NF * TYB * 64);   // [2]: the slot's bytes have landed
    uint64_t* empty = full + 2;                                                  // [2]: every warp has read the slot
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
    FusedM m;
    ftm_init(m, p, lane, ty, cr * TYB + ty, blockIdx.x, blockIdx.y / cl, blockIdx.z, FUN);
    const int nit = m.t.k1 - m.t.k0 + 2;          // planes kp = k0-1 .. k1
    // tensor coordinates of the CTA's first row segment (the maps start at logical (-2, -1, -1))
    const int cx = p.lo[0] + (int)blockIdx.x * FSV_XI;
    const int cy = p.lo[1] + (int)(blockIdx.y / cl) * p.rows_int + cr * TYB;
    const int cz = p.lo[2] + (int)blockIdx.z * p.cz;               // plane k0 - 1
    const bool leader = elect_one();
    const bool feeder = leader && w == 0;
    if (feeder) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], TYB);
        mbar_init(&empty[1], TYB);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ftm_issue<TYB>(maps, ring, &full[0], cx, cy, cz, 0);
        if (nit > 1) ftm_issue<TYB>(maps, ring, &full[1], cx, cy, cz + 1, 1);
    }
    __syncthreads();                               // the barriers exist
    if (cl > 1) fsv_cluster_arrive();
    for (int it = 0; it < nit; ++it) {
        const int kp = m.t.k0 - 1 + it, slot = it & 1;
        d2 sn[FSV_NF];
        // plane it+1 goes where plane it-1 was, as soon as every warp of the CTA has read that slot
        if (feeder && it >= 1 && it + 1 < nit) {
            mbar_wait(&empty[slot ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
            ftm_issue<TYB>(maps, ring, &full[slot ^ 1], cx, cy, cz + it + 1, slot ^ 1);
        }
        mbar_wait(&full[slot], (uint32_t)((it >> 1) & 1));
        ftm_phase_a<TD>(m, p, kp, ring, TYB, slot, it + 1 < nit, sn);
        __syncwarp();                              // every lane of the warp has taken its cells of this slot
        if (leader) mbar_arrive(&empty[slot]);
        // every thread of the cluster has finished reading the exchange buffer that is about to be overwritten, and the
        // stresses of plane kp-1 that phase B reads have been published
        if (cl > 1) fsv_cluster_wait(); else __syncthreads();
        fsv_phase_b<TD, FUN, 0>(m.t, p, kp, sn, TYB, xb, below, rb, above, ra);
        if (cl > 1) fsv_cluster_arrive_relaxed();
    }
    if (cl > 1) fsv_cluster_wait();   // no CTA may exit while a neighbour can still read its shared memory
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
static int g_fuse_tyb = 4, g_fuse_cl = 4, g_fuse_cz = 64, g_fuse_var = 1;   // measured optimum at 767^3 (profiles/)
static bool g_fuse_env = false;
// rows per CTA with an instantiation (profiles/r2_c1_tune_fused_767.log: 2-, 12- and 16-row CTAs and the register-pipelined
// flavour lost on the B200 and were removed)
static bool fsv_tyb_ok(int v) { return v == 4 || v == 6 || v == 8; }

static void fuse_env() {
    if (g_fuse_env) return;
    g_fuse_env = true;
    const char* a = getenv("CHMY_FUSE_TYB");
    const char* b = getenv("CHMY_FUSE_CL");
    const char* c = getenv("CHMY_FUSE_CZ");
    const char* d = getenv("CHMY_FUSE_VARIANT");
    if (d) g_fuse_var = atoi(d);
    if (a) { const int v = atoi(a); if (fsv_tyb_ok(v)) g_fuse_tyb = v; }
    if (b) { const int v = atoi(b); if (v >= 1 && v <= 8) g_fuse_cl = v; }
    if (c) { const int v = atoi(c); if (v >= 1) g_fuse_cz = v; }
}

extern "C" int chmy_set_fused_tuning(int rows_per_cta, int cluster_size, int z_chunk, int variant) {
    fuse_env();
    if (variant >= 0) g_fuse_var = variant;
    if (rows_per_cta > 0) {
        CHMY_REQUIRE(fsv_tyb_ok(rows_per_cta), "rows_per_cta must be 4, 6 or 8");
        g_fuse_tyb = rows_per_cta;
    }
    if (cluster_size > 0) {
        CHMY_REQUIRE(cluster_size >= 1 && cluster_size <= 8, "cluster_size must be 1..8 (the portable cluster limit)");
        g_fuse_cl = cluster_size;
    }
    if (z_chunk > 0) g_fuse_cz = z_chunk;
    return CHMY_OK;
}

template <bool TD, bool FUN, int TYB, int HINT>
static int launch_fused(const FusedP& p, int cl, int variant, dim3 grid, cudaStream_t st) {
    void (*kern)(const FusedP, const int, const int) = k_fused_sv<TD, FUN, TYB, HINT>;
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

typedef CUresult (*chmy_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static chmy_encode_tiled_fn encode_tiled() {
    static chmy_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (chmy_encode_tiled_fn)f;
    }
    return fn;
}

// rank-3 Float64 map over one buffer of a PITCHED field, origin at logical (-2, -1, -1) (16-byte aligned: logical x = 0 sits
// on a 128-byte boundary), box = the CTA's tile of one plane: `rows` segments of 64 cells
static int ftm_make_map(CUtensorMap* map, const chmy_field* f, const double* buf_p0, int rows) {
    chmy_encode_tiled_fn enc = encode_tiled();
    CHMY_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    void* base = (void*)(buf_p0 - 2 - f->stride[1] - f->stride[2]);
    const cuuint64_t dims[3]    = {(cuuint64_t)f->stride[1], (cuuint64_t)f->sd[1], (cuuint64_t)f->sd[2]};
    const cuuint64_t strides[2] = {(cuuint64_t)f->stride[1] * 8, (cuuint64_t)f->stride[2] * 8};
    const cuuint32_t box[3] = {64, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CHMY_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CHMY_OK;
}

template <bool TD, bool FUN, int TYB>
static int launch_fused_tma(const FusedP& p, const FtmMaps& maps, int cl, dim3 grid, cudaStream_t st) {
    void (*kern)(const FusedP, const int, const FtmMaps) = k_fused_tma<TD, FUN, TYB>;
    const size_t smem = ftm_smem_bytes(TYB);
    static bool attr_done[64] = {};
    int dev = 0;
    CHMY_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CHMY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CHMY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
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
    CHMY_CUDA(cudaLaunchKernelEx(&cfg, kern, p, cl, maps));
    return CHMY_OK;
}

template <bool TD, bool FUN, int TYB>
static int launch_fused_hint(const FusedP& p, int cl, int variant, dim3 grid, cudaStream_t st) {
    // cache-policy flavours (variant bits 2..4 -> HINT bits 0..2) exist for the instantiations the headline runs
    if constexpr (!TD && FUN && (TYB == 4 || TYB == 6)) {
        switch ((variant >> 2) & 7) {
        case 1: return launch_fused<TD, FUN, TYB, 1>(p, cl, variant, grid, st);
        case 2: return launch_fused<TD, FUN, TYB, 2>(p, cl, variant, grid, st);
        case 3: return launch_fused<TD, FUN, TYB, 3>(p, cl, variant, grid, st);
        case 6: return launch_fused<TD, FUN, TYB, 6>(p, cl, variant, grid, st);
        case 7: return launch_fused<TD, FUN, TYB, 7>(p, cl, variant, grid, st);
        default: break;
        }
    }
    return launch_fused<TD, FUN, TYB, 0>(p, cl, variant, grid, st);
}

template <bool TD, bool FUN>
static int launch_fused_tyb(const FusedP& p, const FtmMaps* maps, int tyb, int cl, dim3 grid, cudaStream_t st) {
    if (maps) {        // the TMA-fed sweep
        switch (tyb) {
        case 4: return launch_fused_tma<TD, FUN, 4>(p, *maps, cl, grid, st);
        case 6: return launch_fused_tma<TD, FUN, 6>(p, *maps, cl, grid, st);
        default: return launch_fused_tma<TD, FUN, 8>(p, *maps, cl, grid, st);
        }
    }
    switch (tyb) {
    case 4: return launch_fused_hint<TD, FUN, 4>(p, cl, g_fuse_var, grid, st);
    case 6: return launch_fused_hint<TD, FUN, 6>(p, cl, g_fuse_var, grid, st);
    default: return launch_fused_hint<TD, FUN, 8>(p, cl, g_fuse_var, grid, st);
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
int chmy_run_fused(chmy_ctx* ctx, const chmy_launch_desc* ds, const chmy_launch_desc* dv, const Box& box,
                   double* const* cur /* tau[6] Pr V[3] */, double* const* shadow, cudaStream_t st) {
    if (box.n[0] <= 0 || box.n[1] <= 0 || box.n[2] <= 0) return CHMY_OK;
    fuse_env();
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
    p.Gdt = DivC{Gdt, 1.0 / Gdt}; p.eta = DivC{s[0], 1.0 / s[0]}; p.three = DivC{3.0, 1.0 / 3.0};
    p.eve = DivC{s[1], 1.0 / s[1]};
    if (!rho) {
        p.inc.active = 1; p.inc.nd = 3;
        for (int a = 0; a < 3; ++a) {
            p.inc.loc[a] = dv->rho_g.loc[a]; p.inc.origin[a] = dv->grid.origin[a];
            p.inc.spacing[a] = dv->grid.spacing[a]; p.inc.c0[a] = dv->rho_g.c0[a];
        }
        p.inc.r2 = dv->rho_g.r * dv->rho_g.r; p.inc.in = dv->rho_g.in; p.inc.out = dv->rho_g.out;
    }
    const bool td = chmy_force_true_div() || !markstein_ok(Gdt) || !markstein_ok(s[0]) || !markstein_ok(s[1]);
    // geometry: clusters shrink for short boxes (slabs of a split launch)
    int tyb = g_fuse_tyb, cl = g_fuse_cl;
    while (cl > 1 && (cl - 1) * tyb - 2 >= box.n[1]) cl -= 1;
    if (tyb > 4 && cl == 1 && 4 - 2 >= box.n[1]) tyb = 4;
    p.rows_int = cl * tyb - 2;
    const int nch = (box.n[2] + g_fuse_cz - 1) / g_fuse_cz;
    p.cz = (box.n[2] + nch - 1) / nch;
    const dim3 grid((unsigned)((box.n[0] + FSV_XI - 1) / FSV_XI), (unsigned)((box.n[1] + p.rows_int - 1) / p.rows_int * cl),
                    (unsigned)((box.n[2] + p.cz - 1) / p.cz));
    FtmMaps maps;
    const FtmMaps* mp = nullptr;
    if (g_fuse_var & 2) {     // TMA-fed sweep: tensor maps of the eleven ring operands over their CURRENT buffers
        for (int c = 0; c < 6; ++c) CHMY_TRY(ftm_make_map(&maps.m[c], S[c], cur[c], tyb));
        for (int c = 0; c < 5; ++c) CHMY_TRY(ftm_make_map(&maps.m[6 + c], S[11 + c], S[11 + c]->p0, tyb));
        mp = &maps;
    }
    int rc;
    if (rho) rc = td ? launch_fused_tyb<true, false>(p, mp, tyb, cl, grid, st) : launch_fused_tyb<false, false>(p, mp, tyb, cl, grid, st);
    else     rc = td ? launch_fused_tyb<true, true>(p, mp, tyb, cl, grid, st) : launch_fused_tyb<false, true>(p, mp, tyb, cl, grid, st);
    CHMY_TRY(rc);
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
