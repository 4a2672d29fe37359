// api.cu -- C-ABI entry points of libchmy_b200.so: context (device + streams), fields, launch orchestration.
// Interface contract and reference citations: include/chmy_b200.h.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

// declared in bc.cu / ops.cu / ops_fast.cu
int chmy_box_from(const chmy_field* f, const int64_t* lo, const int64_t* hi, Box* out);
int chmy_fill_box(chmy_ctx* ctx, chmy_field* f, double v, const Box& b, cudaStream_t st);
int chmy_copy_box(chmy_ctx* ctx, chmy_field* d, const chmy_field* s, const Box& b, cudaStream_t st);
int chmy_incl_box(chmy_ctx* ctx, chmy_field* f, const InclDev& q, const Box& b, cudaStream_t st);
int chmy_incl_box_f32(chmy_ctx* ctx, chmy_field* f, const InclDevT<float>& q, const Box& b, cudaStream_t st);
int chmy_gauss_box(chmy_ctx* ctx, chmy_field* f, const InclDev& q, const Box& b, cudaStream_t st);
int chmy_gauss_box_f32(chmy_ctx* ctx, chmy_field* f, const InclDevT<float>& q, const Box& b, cudaStream_t st);
int chmy_maxabs_box(chmy_ctx* ctx, const chmy_field* f, const Box& b, unsigned long long* d_out, cudaStream_t st);
int chmy_run_op_generic(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st);
int chmy_run_op_fast(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st, int* handled);

// ---------------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

void chmy_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* chmy_last_error(void) { return g_err; }
extern "C" int chmy_abi_version(void) { return CHMY_ABI_VERSION; }

extern "C" size_t chmy_struct_size(int which) {
    switch (which) {
    case 0: return sizeof(chmy_grid_desc);
    case 1: return sizeof(chmy_batch_desc);
    case 2: return sizeof(chmy_inclusion);
    case 3: return sizeof(chmy_launch_desc);
    case 4: return sizeof(chmy_field_info);
    default: return 0;
    }
}

extern "C" int chmy_device_count(int* count) {
    CHMY_REQUIRE(count != nullptr, "count is NULL");
    CHMY_CUDA(cudaGetDeviceCount(count));
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- context
// live contexts: a Field may outlive its Architecture in a garbage-collected host language (finalizer order is not
// defined), so chmy_field_destroy must not touch a context that is already gone
static std::mutex g_live_mu;
static std::vector<chmy_ctx*> g_live;
static void ctx_register(chmy_ctx* c) {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.push_back(c);
}
static void ctx_unregister(chmy_ctx* c) {
    std::lock_guard<std::mutex> lk(g_live_mu);
    for (size_t i = 0; i < g_live.size(); ++i)
        if (g_live[i] == c) { g_live[i] = g_live.back(); g_live.pop_back(); return; }
}
static bool ctx_alive(const chmy_ctx* c) {
    std::lock_guard<std::mutex> lk(g_live_mu);
    for (chmy_ctx* q : g_live)
        if (q == c) return true;
    return false;
}

static int ctx_init(chmy_ctx* c) {
    CHMY_CUDA(cudaSetDevice(c->device));
    int lo = 0, hi = 0;   // numerically lowest value = highest priority
    CHMY_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // activate!(arch; priority=:high) for the boundary workers (KernelLaunch.jl:44-47)
    CHMY_CUDA(cudaStreamCreateWithPriority(&c->s_main, cudaStreamNonBlocking, lo));
    CHMY_CUDA(cudaStreamCreateWithPriority(&c->s_bnd, cudaStreamNonBlocking, hi));
    CHMY_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CHMY_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CHMY_CUDA(cudaMalloc(&c->d_red, 64 * sizeof(unsigned long long)));
    CHMY_CUDA(cudaMallocHost(&c->h_red, 64 * sizeof(unsigned long long)));
    CHMY_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    CHMY_CUDA(cudaMalloc(&c->d_done, 64 * sizeof(unsigned int)));
    CHMY_CUDA(cudaMemset(c->d_done, 0, 64 * sizeof(unsigned int)));
    chmy_tuning_defaults(&c->tun);
    return CHMY_OK;
}

extern "C" int chmy_ctx_create(int device_id, chmy_ctx** out) {
    CHMY_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    CHMY_CUDA(cudaGetDeviceCount(&ndev));
    // the reference's device ids are 1-based: CuDevice(id - 1)  (ext/ChmyCUDAExt/ChmyCUDAExt.jl:17)
    CHMY_REQUIRE(device_id >= 1 && device_id <= ndev, "device_id %d out of range 1..%d", device_id, ndev);
    chmy_ctx* c = (chmy_ctx*)calloc(1, sizeof(chmy_ctx));
    if (!c) { chmy_set_error("out of host memory"); return CHMY_ERR_NOMEM; }
    c->device = device_id - 1;
    const int rc = ctx_init(c);
    if (rc != CHMY_OK) {           // release whatever was created before the failing call; the error text stays
        if (c->d_done) cudaFree(c->d_done);
        if (c->h_red) cudaFreeHost(c->h_red);
        if (c->d_red) cudaFree(c->d_red);
        if (c->ev_join) cudaEventDestroy(c->ev_join);
        if (c->ev_fork) cudaEventDestroy(c->ev_fork);
        if (c->s_bnd) cudaStreamDestroy(c->s_bnd);
        if (c->s_main) cudaStreamDestroy(c->s_main);
        free(c);
        return rc;
    }
    ctx_register(c);
    *out = c;
    return CHMY_OK;
}

static void upload_release(chmy_ctx* c);     // staged uploads, below

extern "C" int chmy_ctx_destroy(chmy_ctx* c) {
    if (!c || !ctx_alive(c)) return CHMY_OK;
    ctx_unregister(c);
    c->has_pending = 0;          // a deferred launch whose result nobody can observe any more
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm) chmy_comm_destroy(c->comm);
    cudaFree(c->d_red);
    cudaFree(c->d_done);
    cudaFreeHost(c->h_red);
    upload_release(c);
    if (c->ev_time) {
        for (int i = 0; i < CHMY_MAX_EVENTS; ++i) if (c->ev_time[i]) cudaEventDestroy(c->ev_time[i]);
        free(c->ev_time);
    }
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->s_main);
    cudaStreamDestroy(c->s_bnd);
    free(c);
    return CHMY_OK;
}

extern "C" int chmy_ctx_device(const chmy_ctx* c, int* device_id) {
    CHMY_REQUIRE(c && device_id, "NULL argument");
    *device_id = c->device + 1;
    return CHMY_OK;
}

extern "C" int chmy_synchronize(chmy_ctx* c) {
    CHMY_REQUIRE(c != nullptr, "ctx is NULL");
    CHMY_TRY(chmy_flush(c));
    CHMY_CUDA(cudaSetDevice(c->device));
    CHMY_CUDA(cudaStreamSynchronize(c->s_bnd));
    CHMY_CUDA(cudaStreamSynchronize(c->s_main));
    return chmy_comm_check(c->comm);
}

extern "C" int chmy_ctx_launch_count(const chmy_ctx* c, uint64_t* kernels) {
    CHMY_REQUIRE(c && kernels, "NULL argument");
    *kernels = c->n_launches;
    return CHMY_OK;
}

extern "C" int chmy_event_record(chmy_ctx* c, int slot) {
    CHMY_REQUIRE(c && slot >= 0 && slot < CHMY_MAX_EVENTS, "bad event slot");
    CHMY_TRY(chmy_flush(c));
    CHMY_CUDA(cudaSetDevice(c->device));
    if (!c->ev_time) {
        c->ev_time = (cudaEvent_t*)calloc(CHMY_MAX_EVENTS, sizeof(cudaEvent_t));
        if (!c->ev_time) { chmy_set_error("out of host memory"); return CHMY_ERR_NOMEM; }
    }
    if (!c->ev_time[slot]) CHMY_CUDA(cudaEventCreate(&c->ev_time[slot]));
    CHMY_CUDA(cudaEventRecord(c->ev_time[slot], c->s_main));
    return CHMY_OK;
}

// records timing event `slot` on `st` without flushing a deferred launch (used around the fused sweep, ops_fused.cu)
int chmy_event_record_on(chmy_ctx* c, int slot, cudaStream_t st) {
    if (!c->ev_time) {
        c->ev_time = (cudaEvent_t*)calloc(CHMY_MAX_EVENTS, sizeof(cudaEvent_t));
        if (!c->ev_time) { chmy_set_error("out of host memory"); return CHMY_ERR_NOMEM; }
    }
    if (!c->ev_time[slot]) CHMY_CUDA(cudaEventCreate(&c->ev_time[slot]));
    CHMY_CUDA(cudaEventRecord(c->ev_time[slot], st));
    return CHMY_OK;
}

extern "C" int chmy_time_fused_sweep(chmy_ctx* c, int slot_begin, int slot_end) {
    CHMY_REQUIRE(c != nullptr, "ctx is NULL");
    CHMY_REQUIRE((slot_begin < 0 && slot_end < 0) || (slot_begin >= 0 && slot_end >= 0 && slot_begin < CHMY_MAX_EVENTS &&
                 slot_end < CHMY_MAX_EVENTS && slot_begin != slot_end), "bad event slots");
    c->sweep_ev0 = slot_begin < 0 ? 0 : slot_begin + 1;      // stored + 1: a zeroed context times nothing
    c->sweep_ev1 = slot_end < 0 ? 0 : slot_end + 1;
    return CHMY_OK;
}

extern "C" int chmy_event_elapsed_ms(chmy_ctx* c, int a, int b, float* ms) {
    CHMY_REQUIRE(c && ms && c->ev_time && a >= 0 && b >= 0 && a < CHMY_MAX_EVENTS && b < CHMY_MAX_EVENTS &&
                     c->ev_time[a] && c->ev_time[b], "event slots not recorded");
    CHMY_CUDA(cudaEventSynchronize(c->ev_time[b]));
    CHMY_CUDA(cudaEventElapsedTime(ms, c->ev_time[a], c->ev_time[b]));
    return CHMY_OK;
}

extern "C" int chmy_ctx_streams(const chmy_ctx* c, void** main_stream, void** boundary_stream) {
    CHMY_REQUIRE(c != nullptr, "ctx is NULL");
    if (main_stream) *main_stream = (void*)c->s_main;
    if (boundary_stream) *boundary_stream = (void*)c->s_bnd;
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- fields
static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

extern "C" int chmy_field_create(chmy_ctx* ctx, int ndims, const int64_t* dims, const int32_t* loc, int layout,
                                 chmy_field** out) {
    return chmy_field_create_typed(ctx, ndims, dims, loc, layout, CHMY_F64, out);
}

// validates the description and computes the storage layout; no device storage yet
static int field_describe(int ndims, const int64_t* dims, const int32_t* loc, int layout, int dtype, chmy_field** out) {
    CHMY_REQUIRE(dims && loc && out, "NULL argument");
    CHMY_REQUIRE(ndims >= 1 && ndims <= 3, "ndims %d not in 1..3", ndims);
    CHMY_REQUIRE(layout == CHMY_LAYOUT_PITCHED || layout == CHMY_LAYOUT_DENSE, "bad layout %d", layout);
    CHMY_REQUIRE(dtype == CHMY_F64 || dtype == CHMY_F32, "bad element type %d", dtype);
    for (int a = 0; a < ndims; ++a) {
        CHMY_REQUIRE(dims[a] >= 1 && dims[a] < (1ll << 30), "bad field size %lld along dim %d", (long long)dims[a], a + 1);
        CHMY_REQUIRE(loc[a] == CHMY_CENTER || loc[a] == CHMY_VERTEX, "bad location along dim %d", a + 1);
    }
    chmy_field* f = (chmy_field*)calloc(1, sizeof(chmy_field));
    if (!f) { chmy_set_error("out of host memory"); return CHMY_ERR_NOMEM; }
    f->nd = ndims; f->layout = layout;
    f->dtype = dtype; f->esize = dtype == CHMY_F32 ? 4 : 8;
    for (int a = 0; a < 3; ++a) {
        if (a < ndims) {
            f->loc[a] = loc[a]; f->d[a] = dims[a]; f->sd[a] = dims[a] + 4;   // field.jl:58 (halo = 1)
        } else {
            f->loc[a] = CHMY_CENTER; f->d[a] = 1; f->sd[a] = 1;
        }
    }
    // PITCHED: row pitch a multiple of 128 bytes (16 doubles | 32 floats) and a lead-in of one element less, so that
    // logical index 0 of every row sits on a 128-byte boundary (storage element 0 is logical -1).
    const long long per128 = 128 / f->esize;
    const long long pitch = layout == CHMY_LAYOUT_PITCHED ? round_up(f->sd[0], per128) : f->sd[0];
    f->lead      = layout == CHMY_LAYOUT_PITCHED ? per128 - 1 : 0;
    f->stride[0] = 1;
    f->stride[1] = pitch;
    f->stride[2] = pitch * f->sd[1];
    // 64-bit element counts: refuse sizes whose padded volume cannot be a real allocation instead of wrapping around
    const long double volume = (long double)pitch * (long double)f->sd[1] * (long double)f->sd[2];
    if (volume > 1.0e13L) {
        free(f);
        chmy_set_error("field of %lld x %lld x %lld elements is too large", (long long)dims[0], ndims > 1 ? (long long)dims[1] : 1ll,
                       ndims > 2 ? (long long)dims[2] : 1ll);
        return CHMY_ERR_ARG;
    }
    const long long elems = f->lead + pitch * f->sd[1] * f->sd[2] + 2 * per128;
    f->bytes = (size_t)elems * (size_t)f->esize;
    *out = f;
    return CHMY_OK;
}

extern "C" int chmy_field_create_typed(chmy_ctx* ctx, int ndims, const int64_t* dims, const int32_t* loc, int layout,
                                       int dtype, chmy_field** out) {
    CHMY_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    chmy_field* f = nullptr;
    CHMY_TRY(field_describe(ndims, dims, loc, layout, dtype, &f));
    f->ctx = ctx;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaMalloc(&f->alloc, f->bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->alloc, 0, f->bytes, ctx->s_main);   // KernelAbstractions.zeros, field.jl:59
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        chmy_set_error("allocating a field of %zu bytes failed: %s", f->bytes, cudaGetErrorString(e));
        if (f->alloc) cudaFree(f->alloc);
        free(f);
        return CHMY_ERR_NOMEM;
    }
    f->p0 = reinterpret_cast<double*>(reinterpret_cast<char*>(f->alloc) + (size_t)f->esize *
                (size_t)(f->lead + 1 + (ndims > 1 ? f->stride[1] : 0) + (ndims > 2 ? f->stride[2] : 0)));
    *out = f;
    return CHMY_OK;
}

// Descriptor-only field: location, sizes, layout and element type, but no device storage.  chmy_validate_launch accepts
// it; every entry point that would touch storage refuses it.  Lets a binding (and the CPU test-suite) check the
// flattening of `op => args` against the library's own argument rules without a GPU.
extern "C" int chmy_field_create_shell(int ndims, const int64_t* dims, const int32_t* loc, int layout, int dtype,
                                       chmy_field** out) {
    CHMY_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    return field_describe(ndims, dims, loc, layout, dtype, out);
}

extern "C" int chmy_field_destroy(chmy_field* f) {
    if (!f) return CHMY_OK;
    if (ctx_alive(f->ctx)) {     // a deferred launch may still reference the field; kernels in flight may still use it
        chmy_flush(f->ctx);
        cudaSetDevice(f->ctx->device);
        cudaStreamSynchronize(f->ctx->s_bnd);
        cudaStreamSynchronize(f->ctx->s_main);
    }
    if (f->alloc) cudaFree(f->alloc);
    if (f->alt_alloc) cudaFree(f->alt_alloc);
    free(f);
    return CHMY_OK;
}

extern "C" int chmy_field_get_info(const chmy_field* f, chmy_field_info* out) {
    CHMY_REQUIRE(f && out, "NULL argument");
    memset(out, 0, sizeof(*out));
    out->ndims = f->nd; out->layout = f->layout;
    for (int a = 0; a < 3; ++a) { out->loc[a] = f->loc[a]; out->dims[a] = f->d[a]; out->stride[a] = f->stride[a]; }
    if (f->alloc) {        // a descriptor-only field has no addresses
        out->origin_ptr = (void*)f->at_bytes(1, f->nd > 1 ? 1 : 0, f->nd > 2 ? 1 : 0);
        out->base_ptr   = (void*)f->at_bytes(-1, f->nd > 1 ? -1 : 0, f->nd > 2 ? -1 : 0);
    }
    out->bytes      = f->bytes;
    out->dtype      = f->dtype;
    return CHMY_OK;
}

extern "C" int chmy_field_fill(chmy_ctx* ctx, chmy_field* f, double v, const int64_t* lo, const int64_t* hi) {
    CHMY_REQUIRE(ctx && f && lo && hi, "NULL argument");
    Box b;
    CHMY_TRY(chmy_box_from(f, lo, hi, &b));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    return chmy_fill_box(ctx, f, v, b, ctx->s_main);
}

extern "C" int chmy_field_copy(chmy_ctx* ctx, chmy_field* dst, const chmy_field* src, const int64_t* lo, const int64_t* hi) {
    CHMY_REQUIRE(ctx && dst && src && lo && hi, "NULL argument");
    CHMY_REQUIRE(dst->nd == src->nd, "set!(f, other): dimensionality mismatch");
    CHMY_REQUIRE(dst->dtype == src->dtype, "set!(f, other): element types differ");
    Box b, b2;
    CHMY_TRY(chmy_box_from(dst, lo, hi, &b));
    CHMY_TRY(chmy_box_from(src, lo, hi, &b2));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    return chmy_copy_box(ctx, dst, src, b, ctx->s_main);
}

// ---- staged upload.  A strided (3D) host->device copy of a dense host box into the padded field moves 41 GB/s on the B200
// box (rows of a few KB each), a contiguous one 55.6 GB/s (profiles/r2_c35_copy_bandwidth.log): large uploads therefore go
// piecewise and contiguously into one of two device buffers on a copy stream, and a kernel on the main stream scatters each
// piece into the field while the next piece is in flight.  Device->host already runs at the link's rate (56.8 GB/s).
template <class T>
__global__ void __launch_bounds__(256) k_scatter_rows(T* __restrict__ dst, long long s1, long long s2, const T* __restrict__ src, int n0,
                                                       int n1, int j0, int rows) {
    // src: `rows` dense rows of n0 elements, row r = (j, k) with j + k * n1 = j0 + r ; dst: element (0, 0, 0) of the box
    const int r = blockIdx.y;
    if (r >= rows) return;
    const long long q = (long long)j0 + r;
    const long long j = q % n1, k = q / n1;
    T* d = dst + j * s1 + k * s2;
    const T* sr = src + (long long)r * n0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) d[i] = sr[i];
}

static const size_t UP_PIECE = (size_t)64 << 20;

static void upload_release(chmy_ctx* c) {
    for (int b = 0; b < 2; ++b) {
        if (c->d_up[b]) cudaFree(c->d_up[b]);
        if (c->ev_up_copied[b]) cudaEventDestroy(c->ev_up_copied[b]);
        if (c->ev_up_free[b]) cudaEventDestroy(c->ev_up_free[b]);
        c->d_up[b] = nullptr; c->ev_up_copied[b] = nullptr; c->ev_up_free[b] = nullptr;
    }
    if (c->s_up) cudaStreamDestroy(c->s_up);
    c->s_up = nullptr;
    c->d_up_bytes = 0;
}

// false: no staging buffers (out of device memory) -- the caller takes the strided copy
static bool upload_ready(chmy_ctx* c) {
    if (c->d_up_bytes) return true;
    bool ok = cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; b < 2 && ok; ++b)
        ok = cudaMalloc((void**)&c->d_up[b], UP_PIECE) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_up_copied[b], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_up_free[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); upload_release(c); return false; }
    c->d_up_bytes = UP_PIECE;
    return true;
}

template <class T>
static int upload_staged_t(chmy_ctx* ctx, const chmy_field* f, const Box& b, const T* host) {
    const long long n0 = b.n[0], n1 = b.n[1], rows_total = (long long)b.n[1] * b.n[2];
    const long long rows_per_piece = (long long)(UP_PIECE / ((size_t)n0 * sizeof(T)));
    T* dst = reinterpret_cast<T*>(f->at_bytes(b.lo[0], f->nd > 1 ? b.lo[1] : 0, f->nd > 2 ? b.lo[2] : 0));
    int turn = 0;
    for (long long r0 = 0; r0 < rows_total; r0 += rows_per_piece, turn ^= 1) {
        const long long rows = rows_total - r0 < rows_per_piece ? rows_total - r0 : rows_per_piece;
        CHMY_CUDA(cudaStreamWaitEvent(ctx->s_up, ctx->ev_up_free[turn], 0));        // the scatter that read this buffer last
        CHMY_CUDA(cudaMemcpyAsync(ctx->d_up[turn], host + r0 * n0, (size_t)(rows * n0) * sizeof(T), cudaMemcpyHostToDevice, ctx->s_up));
        CHMY_CUDA(cudaEventRecord(ctx->ev_up_copied[turn], ctx->s_up));
        CHMY_CUDA(cudaStreamWaitEvent(ctx->s_main, ctx->ev_up_copied[turn], 0));
        for (long long q0 = 0; q0 < rows; q0 += 32768) {                              // gridDim.y <= 65535
            const int nr = (int)(rows - q0 < 32768 ? rows - q0 : 32768);
            const dim3 grid((unsigned)((n0 + 1023) / 1024 > 0 ? (n0 + 1023) / 1024 : 1), (unsigned)nr);
            k_scatter_rows<T><<<grid, 256, 0, ctx->s_main>>>(dst, (long long)f->stride[1], (long long)f->stride[2],
                                                             reinterpret_cast<const T*>(ctx->d_up[turn]) + q0 * n0, (int)n0, (int)n1,
                                                             (int)(r0 + q0), nr);
            ctx->n_launches++;
        }
        CHMY_CUDA(cudaGetLastError());
        CHMY_CUDA(cudaEventRecord(ctx->ev_up_free[turn], ctx->s_main));
    }
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}

// `host` holds elements of the field's type (Array{Float64} | Array{Float32} in the reference's set!/Array)
static int copy_box_host(chmy_ctx* ctx, const chmy_field* f, void* host, const int64_t* lo, const int64_t* hi, bool to_host) {
    Box b;
    CHMY_TRY(chmy_box_from(f, lo, hi, &b));
    if (b.n[0] <= 0 || b.n[1] <= 0 || b.n[2] <= 0) return CHMY_OK;
    CHMY_TRY(chmy_flush(ctx));
    if (!to_host) const_cast<chmy_field*>(f)->frame_dirty(0);
    CHMY_CUDA(cudaSetDevice(ctx->device));
    // large strided uploads: contiguous pieces + scatter kernel (rows of at most one piece, row index within int range)
    const size_t row_bytes = (size_t)b.n[0] * (size_t)f->esize, box_bytes = row_bytes * (size_t)b.n[1] * (size_t)b.n[2];
    if (!to_host && box_bytes >= ((size_t)32 << 20) && row_bytes <= UP_PIECE && (long long)b.n[1] * b.n[2] < (1ll << 31) &&
        !getenv("CHMY_NO_STAGED_UPLOAD") && upload_ready(ctx)) {
        return f->dtype == CHMY_F64 ? upload_staged_t<double>(ctx, f, b, static_cast<const double*>(host))
                                    : upload_staged_t<float>(ctx, f, b, static_cast<const float*>(host));
    }
    char* dev = f->at_bytes(b.lo[0], f->nd > 1 ? b.lo[1] : 0, f->nd > 2 ? b.lo[2] : 0);
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    const size_t w = (size_t)b.n[0] * (size_t)f->esize;
    cudaPitchedPtr hp = make_cudaPitchedPtr(host, w, (size_t)b.n[0], (size_t)b.n[1]);
    cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)f->stride[1] * (size_t)f->esize, (size_t)f->sd[0], (size_t)f->sd[1]);
    p.srcPtr = to_host ? dp : hp;
    p.dstPtr = to_host ? hp : dp;
    p.extent = make_cudaExtent(w, (size_t)b.n[1], (size_t)b.n[2]);
    p.kind   = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    if (to_host) CHMY_CUDA(cudaStreamSynchronize(ctx->s_bnd));
    CHMY_CUDA(cudaMemcpy3DAsync(&p, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}

extern "C" int chmy_field_copy_from_host(chmy_ctx* ctx, chmy_field* f, const void* src, const int64_t* lo, const int64_t* hi) {
    CHMY_REQUIRE(ctx && f && src && lo && hi, "NULL argument");
    return copy_box_host(ctx, f, const_cast<void*>(src), lo, hi, false);
}

extern "C" int chmy_field_copy_to_host(chmy_ctx* ctx, const chmy_field* f, void* dst, const int64_t* lo, const int64_t* hi) {
    CHMY_REQUIRE(ctx && f && dst && lo && hi, "NULL argument");
    return copy_box_host(ctx, f, dst, lo, hi, true);
}

extern "C" int chmy_host_alloc(chmy_ctx* ctx, size_t bytes, void** out) {
    CHMY_REQUIRE(ctx && out && bytes > 0, "bad argument");
    *out = nullptr;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();      // an allocation failure is not sticky: leave the context usable
        *out = nullptr;
        chmy_set_error("cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return CHMY_ERR_NOMEM;
    }
    return CHMY_OK;
}

extern "C" int chmy_host_free(chmy_ctx* ctx, void* p) {
    if (!p) return CHMY_OK;
    if (ctx && ctx_alive(ctx)) cudaSetDevice(ctx->device);
    CHMY_CUDA(cudaFreeHost(p));
    return CHMY_OK;
}

template <class T>
static InclDevT<T> incl_from(const chmy_grid_desc* g, const chmy_inclusion* inc, const int* loc) {
    InclDevT<T> q;
    memset(&q, 0, sizeof(q));
    q.active = 1; q.nd = g->ndims;
    for (int a = 0; a < 3; ++a) {
        q.loc[a] = loc[a]; q.origin[a] = (T)g->origin[a]; q.spacing[a] = (T)g->spacing[a]; q.c0[a] = (T)inc->c0[a];
    }
    q.r2 = (T)inc->r * (T)inc->r; q.in = (T)inc->in; q.out = (T)inc->out;      // r^2 in the element type
    return q;
}

extern "C" int chmy_field_set_inclusion(chmy_ctx* ctx, chmy_field* f, const chmy_grid_desc* g, const chmy_inclusion* inc) {
    CHMY_REQUIRE(ctx && f && g && inc, "NULL argument");
    CHMY_REQUIRE(f->nd == g->ndims, "set!: field/grid dimensionality mismatch");
    int64_t lo[3] = {1, 1, 1}, hi[3] = {f->d[0], f->d[1], f->d[2]};
    Box b;
    CHMY_TRY(chmy_box_from(f, lo, hi, &b));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    if (f->dtype == CHMY_F32) return chmy_incl_box_f32(ctx, f, incl_from<float>(g, inc, f->loc), b, ctx->s_main);
    return chmy_incl_box(ctx, f, incl_from<double>(g, inc, f->loc), b, ctx->s_main);
}

extern "C" int chmy_field_set_gaussian(chmy_ctx* ctx, chmy_field* f, const chmy_grid_desc* g) {
    CHMY_REQUIRE(ctx && f && g, "NULL argument");
    CHMY_REQUIRE(f->nd == g->ndims, "set!: field/grid dimensionality mismatch");
    int64_t lo[3] = {1, 1, 1}, hi[3] = {f->d[0], f->d[1], f->d[2]};
    Box b;
    CHMY_TRY(chmy_box_from(f, lo, hi, &b));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    chmy_inclusion none;
    memset(&none, 0, sizeof(none));
    if (f->dtype == CHMY_F32) return chmy_gauss_box_f32(ctx, f, incl_from<float>(g, &none, f->loc), b, ctx->s_main);
    return chmy_gauss_box(ctx, f, incl_from<double>(g, &none, f->loc), b, ctx->s_main);
}

extern "C" int chmy_field_maxabs(chmy_ctx* ctx, const chmy_field* f, const int64_t* lo, const int64_t* hi, double* out) {
    CHMY_REQUIRE(ctx && f && lo && hi && out, "NULL argument");
    Box b;
    CHMY_TRY(chmy_box_from(f, lo, hi, &b));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_bnd));
    CHMY_CUDA(cudaMemsetAsync(ctx->d_red, 0, sizeof(unsigned long long), ctx->s_main));
    if (b.n[0] > 0 && b.n[1] > 0 && b.n[2] > 0) CHMY_TRY(chmy_maxabs_box(ctx, f, b, ctx->d_red, ctx->s_main));
    CHMY_CUDA(cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    if (f->dtype == CHMY_F32) {      // the reduction ran on binary32 bit patterns
        const unsigned int bits = (unsigned int)ctx->h_red[0];
        float v;
        memcpy(&v, &bits, sizeof(v));
        *out = (double)v;
    } else {
        memcpy(out, ctx->h_red, sizeof(double));
    }
    return CHMY_OK;
}

// The residual check of the drivers (stokes_3d_inc_ve_T.jl:171-175: maximum(abs.(interior(f))) of four fields, each a
// mapreduce + a host synchronisation in the reference) as ONE round trip: n reductions back to back on the main stream,
// one device-to-host copy, one synchronisation.  lo / hi: n x CHMY_MAX_DIMS logical bounds.
extern "C" int chmy_field_maxabs_many(chmy_ctx* ctx, int n, const chmy_field* const* fields, const int64_t* lo, const int64_t* hi,
                                      double* out) {
    CHMY_REQUIRE(ctx && fields && lo && hi && out, "NULL argument");
    CHMY_REQUIRE(n >= 1 && n <= 64, "between 1 and 64 fields per call, got %d", n);
    Box b[64];
    for (int q = 0; q < n; ++q) {
        CHMY_REQUIRE(fields[q] != nullptr, "field %d is NULL", q);
        CHMY_TRY(chmy_box_from(fields[q], lo + 3 * q, hi + 3 * q, &b[q]));
    }
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_bnd));
    CHMY_CUDA(cudaMemsetAsync(ctx->d_red, 0, (size_t)n * sizeof(unsigned long long), ctx->s_main));
    for (int q = 0; q < n; ++q)
        if (b[q].n[0] > 0 && b[q].n[1] > 0 && b[q].n[2] > 0) CHMY_TRY(chmy_maxabs_box(ctx, fields[q], b[q], ctx->d_red + q, ctx->s_main));
    CHMY_CUDA(cudaMemcpyAsync(ctx->h_red, ctx->d_red, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    for (int q = 0; q < n; ++q) {
        if (fields[q]->dtype == CHMY_F32) {      // the reduction ran on binary32 bit patterns
            const unsigned int bits = (unsigned int)ctx->h_red[q];
            float v;
            memcpy(&v, &bits, sizeof(v));
            out[q] = (double)v;
        } else {
            memcpy(&out[q], &ctx->h_red[q], sizeof(double));
        }
    }
    return CHMY_OK;
}

extern "C" int chmy_halo_slab_len(const chmy_field* f, int dim, int64_t* len) {
    CHMY_REQUIRE(f && len && dim >= 0 && dim < f->nd, "bad argument");
    *len = chmy_slab_len(f, dim);
    return CHMY_OK;
}

static int halo_host(chmy_ctx* ctx, chmy_field* f, int dim, int side, void* host, bool pack) {
    CHMY_REQUIRE(ctx && f && host && dim >= 0 && dim < f->nd && (side == 0 || side == 1), "bad argument");
    const size_t bytes = (size_t)chmy_slab_len(f, dim) * (size_t)f->esize;
    CHMY_TRY(chmy_flush(ctx));
    void* dbuf = nullptr;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaMalloc(&dbuf, bytes));
    int rc = CHMY_OK;
    chmy_field* fs[1] = {f};
    if (pack) {
        rc = chmy_pack_fields(ctx, dim, side, 1, fs, dbuf, ctx->s_main);
        if (rc == CHMY_OK && cudaMemcpyAsync(host, dbuf, bytes, cudaMemcpyDeviceToHost, ctx->s_main) != cudaSuccess) rc = CHMY_ERR_CUDA;
    } else {
        if (cudaMemcpyAsync(dbuf, host, bytes, cudaMemcpyHostToDevice, ctx->s_main) != cudaSuccess) rc = CHMY_ERR_CUDA;
        if (rc == CHMY_OK) rc = chmy_unpack_fields(ctx, dim, side, 1, fs, dbuf, ctx->s_main);
    }
    cudaStreamSynchronize(ctx->s_main);
    cudaFree(dbuf);
    return rc;
}

extern "C" int chmy_halo_pack(chmy_ctx* ctx, const chmy_field* f, int dim, int side, void* host_buf) {
    return halo_host(ctx, const_cast<chmy_field*>(f), dim, side, host_buf, true);
}
extern "C" int chmy_halo_unpack(chmy_ctx* ctx, chmy_field* f, int dim, int side, const void* host_buf) {
    return halo_host(ctx, f, dim, side, const_cast<void*>(host_buf), false);
}

// ---------------------------------------------------------------------------------------------- launch
static int run_op(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st) {
    if (box.n[0] <= 0 || box.n[1] <= 0 || box.n[2] <= 0) return CHMY_OK;
    if (d->op == CHMY_OP_OPERATOR) return chmy_run_op_generic(ctx, d, box, st);   // no tuned variants
    int handled = 0;
    CHMY_TRY(chmy_run_op_fast(ctx, d, box, st, &handled));
    if (handled) return CHMY_OK;
    return chmy_run_op_generic(ctx, d, box, st);
}

static int validate_grid(const chmy_grid_desc* g) {
    CHMY_REQUIRE(g->ndims >= 1 && g->ndims <= 3, "grid.ndims %d not in 1..3", g->ndims);
    for (int a = 0; a < g->ndims; ++a) {
        CHMY_REQUIRE(g->n[a] >= 1 && g->n[a] < (1ll << 30), "bad grid size along dim %d", a + 1);
        for (int s = 0; s < 2; ++s)
            CHMY_REQUIRE(g->connectivity[a][s] == CHMY_BOUNDED || g->connectivity[a][s] == CHMY_CONNECTED,
                         "bad connectivity (only Bounded and Connected are implemented, as in the reference)");
    }
    return CHMY_OK;
}

// bc!(side, dim, ...) for both sides of one dim: FieldBatch sides -> BC kernel, ExchangeBatch sides -> halo exchange
static int bc_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int D, const chmy_batch_desc* l, const chmy_batch_desc* r,
                  cudaStream_t st) {
    CHMY_TRY(chmy_run_bc_dim(ctx, g, D, l, r, st));
    if (l->kind == CHMY_BATCH_EXCHANGE || r->kind == CHMY_BATCH_EXCHANGE) CHMY_TRY(chmy_exchange_dim(ctx, g, D, l, r, st));
    return CHMY_OK;
}

static int validate_batches(const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], bool* any_exchange) {
    *any_exchange = false;
    for (int D = 0; D < g->ndims; ++D)
        for (int s = 0; s < 2; ++s) {
            const chmy_batch_desc& b = bc[D][s];
            CHMY_REQUIRE(b.kind == CHMY_BATCH_EMPTY || b.kind == CHMY_BATCH_FIELD || b.kind == CHMY_BATCH_EXCHANGE, "bad batch kind");
            if (b.kind == CHMY_BATCH_EXCHANGE) {
                // batch_impl(::Connected, ...) is the only producer of ExchangeBatch (batch.jl:98-101)
                CHMY_REQUIRE(g->connectivity[D][s] == CHMY_CONNECTED, "ExchangeBatch on a Bounded side");
                *any_exchange = true;
            }
            CHMY_TRY(chmy_validate_batch(g, D, &b));
        }
    return CHMY_OK;
}

// 64-bit FNV-1a over everything that decides WHICH cells a batch set writes and from what: a field whose frame (the cells
// outside the ops' index range) was last written by batch set X needs no carry-over into its ping-pong twin when the next
// sweep is followed by X again -- X rewrites exactly those cells from cells the sweep has just produced.
static uint64_t batch_signature(const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2]) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    mix(&g->ndims, sizeof(g->ndims));
    for (int D = 0; D < g->ndims; ++D)
        for (int s = 0; s < 2; ++s) {
            const chmy_batch_desc& b = bc[D][s];
            mix(&b.kind, sizeof(b.kind));
            if (b.kind == CHMY_BATCH_EMPTY) continue;
            mix(&b.nfields, sizeof(b.nfields));
            for (int q = 0; q < b.nfields && q < CHMY_MAX_BATCH_FIELDS; ++q) {
                mix(&b.fields[q], sizeof(b.fields[q]));
                if (b.kind != CHMY_BATCH_FIELD) continue;
                mix(&b.bc_kind[q], sizeof(b.bc_kind[q]));
                mix(&b.value[q], sizeof(b.value[q]));
                mix(&b.value_field[q], sizeof(b.value_field[q]));
            }
        }
    return h ? h : 1;
}

// bc!(arch, grid, batchset): D = N..1, side 1 then 2 (batch.jl:20-29)
static int run_batches(chmy_ctx* ctx, const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], cudaStream_t st) {
    ctx->batch_sig = batch_signature(g, bc);
    int handled = 0;
    int rc = ctx->tun.bc_fold ? chmy_run_bc_all(ctx, g, bc, st, &handled) : CHMY_OK;   // every dimension in one launch when nothing is exchanged
    if (!handled)
        for (int D = g->ndims - 1; D >= 0 && rc == CHMY_OK; --D) rc = bc_dim(ctx, g, D, &bc[D][0], &bc[D][1], st);
    ctx->batch_sig = 0;
    return rc;
}

extern "C" int chmy_bc(chmy_ctx* ctx, const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], int flags) {
    CHMY_REQUIRE(ctx && g && bc, "NULL argument");
    CHMY_TRY(validate_grid(g));
    bool any_ex = false;
    CHMY_TRY(validate_batches(g, bc, &any_ex));
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_TRY(run_batches(ctx, g, bc, ctx->s_main));
    if (flags & CHMY_LAUNCH_BLOCKING) CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}

// Launches with boundary batches: 1 (default) = the batches (and the halo exchange) overlap the kernel -- the reference's
// inner region + slabs on two streams for the plain kernels, boundary tiles first + a retire counter for the fused sweeps
// (run_overlapped); 0 = one kernel, then the batches, on one stream.  Results cannot depend on it (outer_width is a hint:
// the ops are pointwise writers).  Deterministic: no run-time tuning, every rank takes the same order.
extern "C" int chmy_set_launch_tuning(chmy_ctx* ctx, int overlap, int bc_fold) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    CHMY_TRY(chmy_flush(ctx));
    if (overlap >= 0) ctx->tun.overlap = overlap > 2 ? 2 : overlap;
    if (bc_fold >= 0) ctx->tun.bc_fold = bc_fold ? 1 : 0;
    return CHMY_OK;
}

extern "C" int chmy_overlapped_count(const chmy_ctx* ctx, uint64_t* n) {
    CHMY_REQUIRE(ctx && n, "NULL argument");
    *n = ctx->n_overlapped;
    return CHMY_OK;
}

// Does a launch with boundary batches split into inner region + slabs, and how wide are the slabs per side?
// outer_width is a scheduling hint (results cannot depend on it: the ops are pointwise writers): unless EXACT_SPLIT is set
// the widths follow the kernel's preference `pref` (or nullptr) and the x widths are nudged so that the inner region and
// the right slab start on even x indices (the tuned kernels own aligned pairs of cells); every cell is still computed
// exactly once.  Pure function of the descriptor and the policy (exported as chmy_launch_split_plan for the CPU tests).
static int plan_split(const chmy_launch_desc* d, const int* pref, int overlap, bool* split_out, int wl[3], int wr[3]) {
    const chmy_grid_desc* g = &d->grid;
    const int N = g->ndims;
    int fulln[3];
    for (int a = 0; a < 3; ++a) fulln[a] = a < N ? (int)g->n[a] + 2 : 1;
    bool any_ex = false;
    CHMY_TRY(validate_batches(g, d->bc, &any_ex));
    bool split = d->has_outer_width != 0;
    if (split) {
        for (int a = 0; a < N; ++a) {
            CHMY_REQUIRE(d->outer_width[a] >= 0, "negative outer_width");
            // the slabs must contain everything the batches touch: halo, first/last interior and send planes
            if (d->outer_width[a] < 3 || 2 * d->outer_width[a] > g->n[a] + 2) split = false;
        }
        // without a neighbour to talk to there is nothing worth six extra launches, so run one full-range kernel
        if (!any_ex && !(d->flags & CHMY_LAUNCH_EXACT_SPLIT)) split = false;
        if (!overlap && !(d->flags & CHMY_LAUNCH_EXACT_SPLIT)) split = false;
    }
    for (int a = 0; a < 3; ++a) wl[a] = wr[a] = 0;
    if (split) {
        for (int a = 0; a < N; ++a) wl[a] = wr[a] = (int)d->outer_width[a];
        if (!(d->flags & CHMY_LAUNCH_EXACT_SPLIT) && pref) {
            for (int a = 0; a < N; ++a)
                if (pref[a] >= 3 && 2 * pref[a] + 2 <= fulln[a]) wl[a] = wr[a] = pref[a];
        }
        if (!(d->flags & CHMY_LAUNCH_EXACT_SPLIT)) {
            wl[0] += wl[0] & 1;
            wr[0] -= (fulln[0] - wr[0]) & 1;            // the right slab starts on an even index; never wider than asked
            if (wr[0] < 3) wr[0] += 2;
            // the nudged widths do not fit (tiny grids): one full-range kernel rather than slabs on odd x origins
            if (wl[0] + wr[0] > fulln[0]) { split = false; for (int a = 0; a < 3; ++a) wl[a] = wr[a] = 0; }
        }
    }
    *split_out = split;
    return CHMY_OK;
}

extern "C" int chmy_launch_split_plan(const chmy_launch_desc* d, const int32_t* pref, int32_t overlap, int32_t* split, int32_t wl[3], int32_t wr[3]) {
    CHMY_REQUIRE(d && split && wl && wr, "NULL argument");
    CHMY_TRY(validate_grid(&d->grid));
    CHMY_REQUIRE(d->has_bc, "a launch without bc is one full-range kernel (KernelLaunch.jl:121-126)");
    bool s = false;
    int l[3], r[3], p[3] = {0, 0, 0};
    if (pref) for (int a = 0; a < 3; ++a) p[a] = pref[a];
    CHMY_TRY(plan_split(d, pref ? p : nullptr, overlap ? 1 : 0, &s, l, r));
    *split = s ? 1 : 0;
    for (int a = 0; a < 3; ++a) { wl[a] = l[a]; wr[a] = r[a]; }
    return CHMY_OK;
}

// Region orchestration of `launch` (KernelLaunch.jl:105-183); RUN(box, stream) executes the op on one region.
// pref: slab widths the op's kernel prefers (outer_width is a hint unless EXACT_SPLIT), or nullptr.
// The split plan is computed by the caller BEFORE anything irreversible (buffer swaps) happens.
template <class RUN>
static int orchestrate(chmy_ctx* ctx, const chmy_launch_desc* d, const RUN& run, bool split, const int wl[3], const int wr[3]) {
    const chmy_grid_desc* g = &d->grid;
    const int N = g->ndims;
    // worksize = ncenters + 2, I = J + Offset(-1)  ->  I in 0..n+1   (KernelLaunch.jl:41,109)
    Box full;
    for (int a = 0; a < 3; ++a) { full.lo[a] = 0; full.n[a] = a < N ? (int)g->n[a] + 2 : 1; }

    if (!d->has_bc) {   // launch_without_bc: one full-range kernel even when the Launcher has an outer_width (:121-126)
        CHMY_TRY(run(full, ctx->s_main));
    } else if (!split) {   // KernelLaunch.jl:156-159
        CHMY_TRY(run(full, ctx->s_main));
        CHMY_TRY(run_batches(ctx, g, d->bc, ctx->s_main));
    } else {        // KernelLaunch.jl:160-181: inner region on the main stream, slabs + batches on the boundary stream
        CHMY_CUDA(cudaEventRecord(ctx->ev_fork, ctx->s_main));
        CHMY_CUDA(cudaStreamWaitEvent(ctx->s_bnd, ctx->ev_fork, 0));
        ctx->batch_sig = batch_signature(g, d->bc);
        for (int D = N - 1; D >= 0; --D) {
            for (int S = 0; S < 2; ++S) {
                Box b;   // outer_worksize / outer_offset, KernelLaunch.jl:63-87
                for (int a = 0; a < 3; ++a) {
                    if (a >= N) { b.lo[a] = 0; b.n[a] = 1; }
                    else if (a < D) { b.lo[a] = 0; b.n[a] = full.n[a]; }
                    else if (a == D) { b.lo[a] = S == 0 ? 0 : full.n[a] - wr[a]; b.n[a] = S == 0 ? wl[a] : wr[a]; }
                    else { b.lo[a] = wl[a]; b.n[a] = full.n[a] - wl[a] - wr[a]; }
                }
                const int rc = run(b, ctx->s_bnd);
                if (rc != CHMY_OK) { ctx->batch_sig = 0; return rc; }
            }
            const int rc = bc_dim(ctx, g, D, &d->bc[D][0], &d->bc[D][1], ctx->s_bnd);
            if (rc != CHMY_OK) { ctx->batch_sig = 0; return rc; }
        }
        ctx->batch_sig = 0;
        Box in;      // inner_worksize / inner_offset, KernelLaunch.jl:60-61
        for (int a = 0; a < 3; ++a) {
            in.lo[a] = a < N ? wl[a] : 0;
            in.n[a]  = a < N ? full.n[a] - wl[a] - wr[a] : 1;
        }
        CHMY_TRY(run(in, ctx->s_main));
        CHMY_CUDA(cudaEventRecord(ctx->ev_join, ctx->s_bnd));
        CHMY_CUDA(cudaStreamWaitEvent(ctx->s_main, ctx->ev_join, 0));
    }
    if (d->flags & CHMY_LAUNCH_BLOCKING) CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));   // KernelLaunch.jl:117
    return CHMY_OK;
}

// ---- lazily fused update_stress! -> update_velocity! (kernel: ops_fused.cu) -------------------
// With fusion enabled, `launch(update_stress!)` (3D, no bc) is deferred; if the next call on the context is the
// matching `launch(update_velocity!; bc)` the two run as ONE sweep, otherwise the deferred launch is executed first
// (every entry point that can observe device state calls chmy_flush), so results never depend on the setting.
extern "C" int chmy_set_fusion(chmy_ctx* ctx, int enable) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    CHMY_TRY(chmy_flush(ctx));
    // bit 0: the 3D stress+velocity sweep; bit 1: additionally the EXPERIMENTAL 2D sweeps (ops_fused2d.cu)
    ctx->fuse = enable ? ((enable & 3) ? (enable & 3) : 1) : 0;
    return CHMY_OK;
}

extern "C" int chmy_fused_count(const chmy_ctx* ctx, uint64_t* sweeps) {
    CHMY_REQUIRE(ctx && sweeps, "NULL argument");
    *sweeps = ctx->n_fused;
    return CHMY_OK;
}

extern "C" int chmy_fusion_fallback_count(const chmy_ctx* ctx, uint64_t* pairs) {
    CHMY_REQUIRE(ctx && pairs, "NULL argument");
    *pairs = ctx->n_fuse_fallback;
    return CHMY_OK;
}

static int run_plain(chmy_ctx* ctx, const chmy_launch_desc* d) {
    bool split = false;
    int wl[3] = {0, 0, 0}, wr[3] = {0, 0, 0};
    if (d->has_bc) CHMY_TRY(plan_split(d, nullptr, ctx->tun.overlap, &split, wl, wr));
    return orchestrate(ctx, d, [&](const Box& b, cudaStream_t st) { return run_op(ctx, d, b, st); }, split, wl, wr);
}

int chmy_flush(chmy_ctx* ctx) {
    if (!ctx || !ctx->has_pending) return CHMY_OK;
    ctx->has_pending = 0;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    return run_plain(ctx, &ctx->pending);
}

static int ensure_shadow(chmy_ctx* ctx, chmy_field* f) {
    if (f->alt_alloc) return CHMY_OK;
    cudaError_t e = cudaMalloc(&f->alt_alloc, f->bytes);
    if (e != cudaSuccess) {
        f->alt_alloc = nullptr;
        cudaGetLastError();
        chmy_set_error("cudaMalloc of a %zu-byte shadow buffer failed: %s", f->bytes, cudaGetErrorString(e));
        return CHMY_ERR_NOMEM;
    }
    CHMY_CUDA(cudaMemsetAsync(f->alt_alloc, 0, f->bytes, ctx->s_main));
    f->frame_synced = false;
    f->frame_writer = 0;
    return CHMY_OK;
}

// Which ping-pong fields need their frame (cells outside the ops' index range, never produced by a sweep) carried over into
// the shadow buffer: those whose frame changed since the buffers last agreed -- unless the only writer was the very batch
// set `sig` that follows this sweep too (it rewrites those cells in the new buffer before anything reads them).
static int frames_to_carry(chmy_field* const* pp, int npp, uint64_t sig, chmy_field** fr, double** fsrc, double** fdst) {
    int nfr = 0;
    for (int q = 0; q < npp; ++q) {
        chmy_field* f = pp[q];
        if (f->frame_synced) continue;
        if (sig != 0 && f->frame_writer == sig) continue;
        fr[nfr] = f; fsrc[nfr] = f->p0; fdst[nfr] = f->alt_p0(); ++nfr;
        f->frame_synced = true; f->frame_writer = 0;
    }
    return nfr;
}

// cuStreamWaitValue32: the boundary stream sleeps on the sweep's retire counter without occupying an SM
typedef int (*chmy_wait_value_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static chmy_wait_value_fn wait_value_fn() {
    static chmy_wait_value_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (chmy_wait_value_fn)f;
        (void)cudaGetLastError();
    }
    return fn;
}
int chmy_spin_until(chmy_ctx* ctx, const unsigned int* counter, unsigned int target, cudaStream_t st);   // bc.cu: one-thread fallback

// One fused 3D sweep whose boundary batches overlap it WITHOUT splitting the launch: the sweep runs its boundary tiles
// first and counts them as they retire; the boundary stream (highest priority) carries the frame over, sleeps on that
// counter, then applies the batches / packs, exchanges and unpacks the halos while the interior tiles are still running.
// The next launch on the main stream waits for the boundary stream (KernelLaunch.jl:160-181 semantics, one launch).
static int run_overlapped(chmy_ctx* ctx, const chmy_launch_desc* ds, const chmy_launch_desc* dv, double* const* cur,
                          double* const* shadow, int nfr, chmy_field* const* fr, double* const* fsrc, double* const* fdst) {
    const chmy_grid_desc* g = &dv->grid;
    Box full;
    for (int a = 0; a < 3; ++a) { full.lo[a] = 0; full.n[a] = (int)g->n[a] + 2; }
    unsigned int* done = ctx->d_done + (ctx->n_overlapped & 31);      // a fresh word per launch (zeroed below)
    CHMY_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned int), ctx->s_main));
    CHMY_CUDA(cudaEventRecord(ctx->ev_fork, ctx->s_main));
    unsigned int target = 0;
    CHMY_TRY(chmy_run_fused(ctx, ds, dv, full, cur, shadow, ctx->s_main, done, &target));
    CHMY_CUDA(cudaStreamWaitEvent(ctx->s_bnd, ctx->ev_fork, 0));
    CHMY_TRY(chmy_frame_copy(ctx, &ds->grid, nfr, fr, fsrc, fdst, ctx->s_bnd));   // frame cells: untouched by the sweep
    if (chmy_wait_value_fn wv = wait_value_fn()) {
        const int rc = wv(ctx->s_bnd, (unsigned long long)(uintptr_t)done, target, 0x1u /* CU_STREAM_WAIT_VALUE_GEQ */);
        CHMY_REQUIRE(rc == 0, "cuStreamWaitValue32 failed (%d)", rc);
    } else {
        CHMY_TRY(chmy_spin_until(ctx, done, target, ctx->s_bnd));
    }
    CHMY_TRY(run_batches(ctx, g, dv->bc, ctx->s_bnd));
    CHMY_CUDA(cudaEventRecord(ctx->ev_join, ctx->s_bnd));
    CHMY_CUDA(cudaStreamWaitEvent(ctx->s_main, ctx->ev_join, 0));
    ctx->n_overlapped++;
    if (dv->flags & CHMY_LAUNCH_BLOCKING) CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));   // KernelLaunch.jl:117
    return CHMY_OK;
}

static int run_fused(chmy_ctx* ctx, const chmy_launch_desc* ds, const chmy_launch_desc* dv) {
    // everything that can fail for reasons of the descriptor is decided BEFORE the buffers are swapped
    const bool exact = (dv->flags & CHMY_LAUNCH_EXACT_SPLIT) != 0;
    bool split = false;
    int wl[3] = {0, 0, 0}, wr[3] = {0, 0, 0};
    // slabs of a literal split follow outer_width; otherwise the overlap needs no slabs at all (run_overlapped)
    bool any_ex = false;
    if (dv->has_bc) CHMY_TRY(validate_batches(&dv->grid, dv->bc, &any_ex));
    if (dv->has_bc && exact) CHMY_TRY(plan_split(dv, nullptr, 1, &split, wl, wr));
    if (split) CHMY_REQUIRE((wl[0] & 1) == 0 && ((dv->grid.n[0] + 2 - wr[0]) & 1) == 0, "fused sweep: x slabs must start on even indices");
    // ping-pong fields: tau[6], Pr, V[3]
    chmy_field* pp[10];
    for (int c = 0; c < 6; ++c) pp[c] = ds->fields[c];
    pp[6] = ds->fields[6];
    for (int c = 0; c < 3; ++c) pp[7 + c] = dv->fields[c];
    for (int q = 0; q < 10; ++q)
        if (ensure_shadow(ctx, pp[q]) != CHMY_OK) return 1;      // out of memory: the caller falls back to two kernels
    // cells outside [0, n+1]^3 are not produced by the sweep: carry them over where they may have changed
    const uint64_t sig = dv->has_bc ? batch_signature(&dv->grid, dv->bc) : 0;
    chmy_field* fr[10];
    double *fsrc[10], *fdst[10];
    const int nfr = frames_to_carry(pp, 10, sig, fr, fsrc, fdst);
    double *cur[10], *shadow[10];
    for (int q = 0; q < 10; ++q) { cur[q] = pp[q]->p0; shadow[q] = pp[q]->alt_p0(); }
    // from here on the fields ARE their new buffers: the boundary batches of this launch act on the new V
    for (int q = 0; q < 10; ++q) pp[q]->swap_buffers();
    ctx->n_fused++;
    // without a neighbour the batches are one small launch: nothing worth hiding (measured: 20.41 vs 20.44 ms at 767^3)
    if (dv->has_bc && !exact && any_ex && ctx->tun.overlap >= 1) return run_overlapped(ctx, ds, dv, cur, shadow, nfr, fr, fsrc, fdst);
    if (dv->has_bc && !exact && ctx->tun.overlap == 2) return run_overlapped(ctx, ds, dv, cur, shadow, nfr, fr, fsrc, fdst);
    CHMY_TRY(chmy_frame_copy(ctx, &ds->grid, nfr, fr, fsrc, fdst, ctx->s_main));
    return orchestrate(ctx, dv, [&](const Box& b, cudaStream_t st) { return chmy_run_fused(ctx, ds, dv, b, cur, shadow, st); }, split, wl, wr);
}

// 2D sweeps and the 3D thermal sweep (ops_fused2d.cu): same protocol as run_fused -- shadow buffers, frame carry-over, swap,
// then the usual region orchestration with the sweep as the region kernel.  Returns 1 when the shadow buffers cannot be
// allocated.
static int run_fused2d(chmy_ctx* ctx, int kind, const chmy_launch_desc* dp, const chmy_launch_desc* dc) {
    // x slabs of a split launch: one row segment (60 interior cells in 2D, 64 cells in the 3D thermal sweep); y, z as asked
    const int pref[3] = {kind == 4 ? 64 : 60, 0, 0};
    bool split = false;
    int wl[3] = {0, 0, 0}, wr[3] = {0, 0, 0};
    if (dc->has_bc) CHMY_TRY(plan_split(dc, pref, ctx->tun.overlap, &split, wl, wr));
    if (split) CHMY_REQUIRE((wl[0] & 1) == 0 && ((dc->grid.n[0] + 2 - wr[0]) & 1) == 0, "fused sweep: x slabs must start on even indices");
    chmy_field* pp[6];
    const int npp = chmy_fused2d_pingpong(kind, dp, dc, pp);
    for (int q = 0; q < npp; ++q)
        if (ensure_shadow(ctx, pp[q]) != CHMY_OK) return 1;
    const uint64_t sig = dc->has_bc ? batch_signature(&dc->grid, dc->bc) : 0;
    chmy_field* fr[6];
    double *fsrc[6], *fdst[6];
    const int nfr = frames_to_carry(pp, npp, sig, fr, fsrc, fdst);
    if (dp->grid.ndims == 3) CHMY_TRY(chmy_frame_copy(ctx, &dp->grid, nfr, fr, fsrc, fdst, ctx->s_main));
    else CHMY_TRY(chmy_frame_copy2(ctx, &dp->grid, nfr, fr, fsrc, fdst, ctx->s_main));
    double *cur[6], *shadow[6];
    for (int q = 0; q < npp; ++q) { cur[q] = pp[q]->p0; shadow[q] = pp[q]->alt_p0(); }
    for (int q = 0; q < npp; ++q) pp[q]->swap_buffers();
    ctx->n_fused++;
    return orchestrate(ctx, dc, [&](const Box& b, cudaStream_t st) { return chmy_run_fused2d(ctx, kind, dp, dc, b, cur, shadow, st); }, split, wl, wr);
}

// The argument checks of chmy_launch, without a device (fields may be descriptor-only, chmy_field_create_shell).
extern "C" int chmy_validate_launch(const chmy_launch_desc* d) {
    CHMY_REQUIRE(d != nullptr, "NULL argument");
    CHMY_TRY(validate_grid(&d->grid));
    CHMY_REQUIRE(d->nfields >= 0 && d->nfields <= CHMY_MAX_OP_FIELDS && d->nscalars >= 0 && d->nscalars <= CHMY_MAX_SCALARS,
                 "bad field / scalar count");
    CHMY_TRY(chmy_validate_op(d));
    CHMY_REQUIRE(d->op != CHMY_OP_NONE, "chmy_launch needs an op (use chmy_bc for a bare batch set)");
    if (d->has_bc) {
        bool any_ex = false;
        CHMY_TRY(validate_batches(&d->grid, d->bc, &any_ex));
    }
    if (d->has_outer_width)
        for (int a = 0; a < d->grid.ndims; ++a) CHMY_REQUIRE(d->outer_width[a] >= 0, "negative outer_width");
    return CHMY_OK;
}

extern "C" int chmy_launch(chmy_ctx* ctx, const chmy_launch_desc* d) {
    CHMY_REQUIRE(ctx && d, "NULL argument");
    const chmy_grid_desc* g = &d->grid;
    CHMY_TRY(chmy_validate_launch(d));
    for (int q = 0; q < d->nfields; ++q)
        CHMY_REQUIRE(!d->fields[q] || d->fields[q]->alloc, "field %d is descriptor-only (chmy_field_create_shell): it has no storage", q);
    CHMY_CUDA(cudaSetDevice(ctx->device));
    // a literal split (EXACT_SPLIT) whose x slabs do not start on even indices cannot feed the sweep's aligned cell pairs
    const bool odd_exact_split = (d->flags & CHMY_LAUNCH_EXACT_SPLIT) && d->has_bc && d->has_outer_width &&
                                 ((d->outer_width[0] & 1) || ((g->n[0] + 2 - d->outer_width[0]) & 1));
    if (ctx->has_pending && d->op == CHMY_OP_UPDATE_VELOCITY && chmy_fused_eligible(&ctx->pending, d) && !odd_exact_split) {
        ctx->has_pending = 0;
        const int rc = run_fused(ctx, &ctx->pending, d);
        if (rc <= 0) return rc;
        ctx->n_fuse_fallback++;
        CHMY_TRY(run_plain(ctx, &ctx->pending));    // no memory for the shadow buffers: two kernels
        return run_plain(ctx, d);
    }
    if (ctx->has_pending && (ctx->fuse & 2) && !odd_exact_split) {       // EXPERIMENTAL pairs: 2D kinds 1-3, 3D thermal kind 4
        const int kind = chmy_fused2d_kind(&ctx->pending, d);
        if (kind) {
            ctx->has_pending = 0;
            const int rc = run_fused2d(ctx, kind, &ctx->pending, d);
            if (rc <= 0) return rc;
            ctx->n_fuse_fallback++;
            CHMY_TRY(run_plain(ctx, &ctx->pending));    // no memory for the shadow buffers: two kernels
            return run_plain(ctx, d);
        }
    }
    CHMY_TRY(chmy_flush(ctx));
    if ((ctx->fuse & 1) && d->op == CHMY_OP_UPDATE_STRESS && g->ndims == 3 && !d->has_bc) {
        ctx->pending = *d;
        ctx->has_pending = 1;
        return CHMY_OK;
    }
    if ((ctx->fuse & 2) && !d->has_bc &&
        ((g->ndims == 2 && (d->op == CHMY_OP_UPDATE_STRESS || d->op == CHMY_OP_COMPUTE_Q || d->op == CHMY_OP_UPDATE_THERMAL_FLUX)) ||
         (g->ndims == 3 && d->op == CHMY_OP_UPDATE_THERMAL_FLUX))) {
        ctx->pending = *d;
        ctx->has_pending = 1;
        return CHMY_OK;
    }
    return run_plain(ctx, d);
}
