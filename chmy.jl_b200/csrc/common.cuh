// common.cuh -- internal types shared by the translation units of libchmy_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/chmy_b200.h"

// ---------------------------------------------------------------------------------------------- errors
void chmy_set_error(const char* fmt, ...);

#define CHMY_CUDA(call)                                                                               \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            chmy_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,    \
                           cudaGetErrorString(_e));                                                   \
            return CHMY_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

#define CHMY_REQUIRE(cond, ...)                                                                       \
    do {                                                                                              \
        if (!(cond)) {                                                                                \
            chmy_set_error(__VA_ARGS__);                                                              \
            return CHMY_ERR_ARG;                                                                      \
        }                                                                                             \
    } while (0)

#define CHMY_TRY(expr)                                                                                \
    do {                                                                                              \
        int _s = (expr);                                                                              \
        if (_s != CHMY_OK) return _s;                                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------- device views

// View of a Field on the device.  p addresses logical index 0 of every active dimension
// (reference indexing: f[I] = data[I + 2H], src/Fields/field.jl:18); x stride is 1.
// The solvers are Float64 programs (FV); the element-type-generic pieces (fields, bc!, halo slabs, grid operators)
// are instantiated for Float32 as well, as the reference's tests are (test/common.jl:9).
template <class T>
struct FVT {
    T* __restrict__ p;
    long long sy, sz;   // element strides of dims 2 and 3 (0 when inactive)
};
using FV = FVT<double>;

template <class T>
__device__ __forceinline__ T fv_ld(const FVT<T>& f, int i, int j, int k) {
    return f.p[(long long)i + (long long)j * f.sy + (long long)k * f.sz];
}
template <class T>
__device__ __forceinline__ void fv_st(const FVT<T>& f, int i, int j, int k, T v) {
    f.p[(long long)i + (long long)j * f.sy + (long long)k * f.sz] = v;
}

struct Box {
    int lo[3];
    int n[3];   // extents (>= 1 for inactive dims)
};

// FunctionField `init_incl` evaluated in-kernel (function_field.jl:49-59; uniform_axis.jl:18-19).
template <class T>
struct InclDevT {
    int active;
    int nd;
    int loc[3];
    T   origin[3], spacing[3], c0[3];
    T   r2, in, out;
};
using InclDev = InclDevT<double>;

template <class T>
__device__ __forceinline__ T coord_dev(T origin, T spacing, int loc, int i) {
    const T im1 = (T)(i - 1);
    return loc == CHMY_VERTEX ? fma(im1, spacing, origin) : fma(im1, spacing, fma((T)0.5, spacing, origin));
}

template <class T>
__device__ __forceinline__ T incl_eval(const InclDevT<T>& q, int i, int j, int k) {
    const int I[3] = {i, j, k};
    T s = (T)0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d < q.nd) {
            const T c  = coord_dev(q.origin[d], q.spacing[d], q.loc[d], I[d]) - q.c0[d];
            const T c2 = c * c;
            s = (d == 0) ? c2 : s + c2;
        }
    }
    return s < q.r2 ? q.in : q.out;
}

// Julia Base.max/min on Float64 (NaN-propagating, max(-0.0,+0.0) = +0.0)
__device__ __forceinline__ double jl_max0(double v) {
    return (v != v) ? v : fmax(v, 0.0);
}
__device__ __forceinline__ double jl_min0(double v) {
    return (v != v) ? v : fmin(v, 0.0);
}

// ---------------------------------------------------------------------------------------------- host objects

struct chmy_field {
    chmy_ctx* ctx;
    int       nd;
    int       layout;
    int       dtype;       // chmy_dtype; strides, lead and offsets below count ELEMENTS of that type
    int       esize;       // 8 | 4
    int       loc[3];
    long long d[3];        // logical dims (1 for inactive)
    long long sd[3];       // logical storage dims d+4 (1 for inactive)
    long long stride[3];   // element strides (stride[0] = 1)
    long long lead;        // elements between the allocation base and storage element (-1,-1,-1)
    double*   alloc;       // cudaMalloc'ed base
    size_t    bytes;
    double*   p0;          // address of logical (0,0,0) over active dims (a float* in disguise for CHMY_F32 fields:
                           // only viewT<float>() / at_bytes() may touch those)
    // ping-pong shadow of the fused stress+velocity sweep (ops_fused.cu): lazily allocated twin of `alloc`; the two
    // are swapped after every fused launch.  frame_synced: the cells outside the ops' index range [0, n+1]^N hold the
    // same values in both buffers (cleared by everything that may write such cells: fill/copy/set!/bc!/halo unpack).
    double*   alt_alloc;
    bool      frame_synced;
    // who wrote the frame since the buffers last agreed: the signature of ONE boundary-batch set (its cells are rewritten
    // by the same batch set after the next sweep, so they need no carry-over), or 0 = anything else / several writers
    uint64_t  frame_writer;
    void      frame_dirty(uint64_t sig) {
        if (frame_synced) frame_writer = sig;
        else if (frame_writer != sig) frame_writer = 0;
        frame_synced = false;
    }
    double*   alt_p0() const { return alt_alloc + (p0 - alloc); }
    void      swap_buffers() { double* a = alloc; const ptrdiff_t o = p0 - alloc; alloc = alt_alloc; alt_alloc = a; p0 = alloc + o; }

    FV view() const { return FV{p0, nd > 1 ? stride[1] : 0, nd > 2 ? stride[2] : 0}; }      // Float64 fields only
    template <class T>
    FVT<T> viewT() const { return FVT<T>{reinterpret_cast<T*>(p0), nd > 1 ? stride[1] : 0, nd > 2 ? stride[2] : 0}; }
    double* at(long long i, long long j, long long k) const {                                // Float64 fields only
        return p0 + i + (nd > 1 ? j * stride[1] : 0) + (nd > 2 ? k * stride[2] : 0);
    }
    char* at_bytes(long long i, long long j, long long k) const {
        return reinterpret_cast<char*>(p0) + (size_t)esize * (size_t)(i + (nd > 1 ? j * stride[1] : 0) + (nd > 2 ? k * stride[2] : 0));
    }
};

struct chmy_comm;   // comm.cu
int chmy_comm_check(const chmy_comm* c);   // comm.cu: has a peer-store flag wait timed out?

// Per-context tuning (set from the environment when the context is created; chmy_set_fused_tuning & co. change it)
struct chmy_tuning {
    int fuse_tyb, fuse_cl, fuse_cz, fuse_var;   // fused 3D sweep: rows per CTA, CTAs per cluster, planes per z-chunk, barrier flavour
    int f2_cy, f2_unroll, t3_cz;                // 2D sweeps: rows per y-chunk, rows per load group; 3D thermal sweep: planes per chunk
    int overlap;                                // launches with boundary batches: 1 = overlap them with the kernel (default), 0 = one stream
    int bc_fold;                                // 1 (default): a batch set without exchange runs as ONE launch (k_bc_all); 0: one per dimension
};

struct chmy_ctx {
    int          device;        // 0-based CUDA ordinal
    cudaStream_t s_main;        // inner-domain work
    cudaStream_t s_bnd;         // boundary slabs, BC, pack/exchange/unpack (highest priority)
    cudaEvent_t  ev_fork, ev_join;
    unsigned long long* d_red;  // device scratch for reductions (8 slots)
    unsigned long long* h_red;  // pinned host mirror
    double*      h_stage;       // pinned staging for small host transfers
    size_t       h_stage_bytes;
    uint64_t     n_launches;
    int          sm_count;
    chmy_comm*   comm;          // null on a single-device architecture
    cudaEvent_t* ev_time;       // lazily created timing events (CHMY_MAX_EVENTS slots)
    // lazily fused update_stress! -> update_velocity! (api.cu): a deferred stress launch waiting for its velocity launch
    int               fuse;          // chmy_set_fusion: bit 0 = 3D stress+velocity sweep, bit 1 = + experimental 2D sweeps
    int               has_pending;
    chmy_launch_desc  pending;
    uint64_t          n_fused;       // fused sweeps launched so far
    uint64_t          n_fuse_fallback;   // pairs that had to run as two kernels because their shadow buffers could not be allocated
    chmy_tuning       tun;
    unsigned int*     d_done;        // device counter of the boundary-first sweep (ops_fused.cu): CTAs of boundary tiles retired
    uint64_t          batch_sig;     // signature of the batch set being applied right now (0 outside launch / bc!)
    uint64_t          n_overlapped;  // launches whose batches ran behind the boundary tiles of a still-running sweep
    // fused 3D sweep: the division mode of the last launch (0 four operations, 1 div.rn.f64, 2 two operations) and a small
    // cache of div2_exact() verdicts (fast_common.cuh)
    int               div_mode, n_div2, div2_next;
    double            div2_c[8];
    bool              div2_ok[8];
    int               sweep_ev0, sweep_ev1;   // chmy_time_fused_sweep: event slots (+1) recorded right before / after the 3D sweep kernel, 0 = off
    // staged uploads (api.cu copy_box_host): two device buffers a dense host box is copied into piecewise, contiguously, while
    // a kernel scatters the previous piece into the padded field
    char*             d_up[2];
    size_t            d_up_bytes;
    cudaStream_t      s_up;
    cudaEvent_t       ev_up_copied[2], ev_up_free[2];
};

// api.cu: runs a deferred update_stress! launch now (every entry point that reads or writes device state calls it)
int chmy_flush(chmy_ctx* ctx);

static inline dim3 grid_for(const Box& b, dim3 blk) {
    return dim3((unsigned)((b.n[0] + blk.x - 1) / blk.x), (unsigned)((b.n[1] + blk.y - 1) / blk.y),
                (unsigned)((b.n[2] + blk.z - 1) / blk.z));
}

// ops_fused.cu: div2_exact() (fast_common.cuh) per divisor, remembered per context
bool chmy_div2_cached(chmy_ctx* ctx, double c);
// api.cu
int chmy_event_record_on(chmy_ctx* c, int slot, cudaStream_t st);
// ops.cu
int chmy_run_op(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st);
int chmy_validate_op(const chmy_launch_desc* d);
// bc.cu
int chmy_validate_batch(const chmy_grid_desc* g, int dim, const chmy_batch_desc* b);
int chmy_run_bc_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int dim, const chmy_batch_desc* left,
                    const chmy_batch_desc* right, cudaStream_t st);
int chmy_run_bc_all(chmy_ctx* ctx, const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], cudaStream_t st, int* handled);
// comm.cu
int chmy_comm_destroy(chmy_comm* c);
int chmy_exchange_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int dim, const chmy_batch_desc* left,
                      const chmy_batch_desc* right, cudaStream_t st);
// ops_fused.cu
bool chmy_fused_eligible(const chmy_launch_desc* ds, const chmy_launch_desc* dv);
// done != nullptr: boundary tiles first, each of their CTAs adds 1 to *done when retired; *n_signal = how many will
int chmy_run_fused(chmy_ctx* ctx, const chmy_launch_desc* ds, const chmy_launch_desc* dv, const Box& box,
                   double* const* cur, double* const* shadow, cudaStream_t st, unsigned int* done = nullptr,
                   unsigned int* n_signal = nullptr);
void chmy_tuning_defaults(chmy_tuning* t);
int chmy_frame_copy(chmy_ctx* ctx, const chmy_grid_desc* g, int n, chmy_field* const* fs, double* const* src,
                    double* const* dst, cudaStream_t st);
// ops_fused2d.cu (EXPERIMENTAL 2D sweeps)
int chmy_fused2d_kind(const chmy_launch_desc* dp, const chmy_launch_desc* dc);
int chmy_fused2d_pingpong(int kind, const chmy_launch_desc* dp, const chmy_launch_desc* dc, chmy_field** pp);
int chmy_run_fused2d(chmy_ctx* ctx, int kind, const chmy_launch_desc* dp, const chmy_launch_desc* dc, const Box& box,
                     double* const* cur, double* const* shadow, cudaStream_t st);
int chmy_frame_copy2(chmy_ctx* ctx, const chmy_grid_desc* g, int n, chmy_field* const* fs, double* const* src,
                     double* const* dst, cudaStream_t st);
// bc.cu (halo slabs)
// dbuf holds the slabs in the fields' element type (all fields of one call share it)
int chmy_pack_fields(chmy_ctx* ctx, int dim, int side, int nf, chmy_field* const* fs, void* dbuf, cudaStream_t st);
int chmy_unpack_fields(chmy_ctx* ctx, int dim, int side, int nf, chmy_field* const* fs, const void* dbuf,
                       cudaStream_t st);
long long chmy_slab_len(const chmy_field* f, int dim);
