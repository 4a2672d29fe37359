// comm.cu -- CartesianTopology and the distributed halo exchange (one process per GPU, NCCL over NVLink/NVSwitch).
//
// Replaces src/Distributed/{topology.jl,exchange_halo.jl,stack_allocator.jl,task_local_exchanger.jl}: the
// reference posts one MPI Irecv/Isend pair per field per (dim, side) with a host busy-poll; here the slabs of all
// fields of a (dim, side) are packed by one kernel into one message, both sides of a dimension travel in a single
// NCCL group on the boundary stream, and nothing synchronises with the host.
//
// NCCL is resolved at run time with dlopen so that a single-device build has no NCCL dependency and a Julia or
// Python host process that already carries a libnccl.so.2 shares it.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; every call goes through the table below
#include <stdlib.h>

#include "common.cuh"
#include "peer_link.cuh"   // protocol of the peer-store exchange (CHMY_EXCHANGE_PEER), shared with its host emulation

struct NcclApi {
    void* handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char* (*GetErrorString)(ncclResult_t);
};

static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return CHMY_OK;
    const char* names[] = {getenv("CHMY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { chmy_set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return CHMY_ERR_NCCL; }
#define SYM(field, name)                                                                   \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                             \
    if (!g_nccl.field) { chmy_set_error("NCCL symbol %s missing", name); return CHMY_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return CHMY_OK;
}

#define CHMY_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t _r = (call);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            chmy_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
            return CHMY_ERR_NCCL;                                                                    \
        }                                                                                            \
    } while (0)

struct chmy_comm {
    int        nranks, rank, nd;
    int        dims[3], coords[3];
    int        nb[3][2];            // -1 == MPI.PROC_NULL
    ncclComm_t nccl;
    double*    sbuf[2];             // per side: packed send slabs of all fields
    double*    rbuf[2];
    size_t     cap[2];              // elements
    double*    d_scal;              // scratch for scalar all-reduces
    // ---- peer-store exchange (CHMY_EXCHANGE_PEER; peer_link.cuh): one link per neighbour
    int        xmode;               // chmy_exchange_mode
    PlLink     link[3][2];
    void*      grave[64];           // blocks replaced by larger ones or never paired: the peer may still have them mapped,
    int        ngrave;              //   so they are only freed with the communicator
    char*      d_stage;             // hand-shake staging: 128 B out, 128 B in
    char*      h_stage;             // pinned twin
    int*       h_err;               // pinned + mapped: non-zero once a flag wait has timed out
    int*       d_err;               // device alias of h_err
    unsigned long long timeout_ns;  // flag waits give up after this long instead of hanging the device
    uint64_t   n_peer_msgs, n_nccl_msgs;
};

// ---------------------------------------------------------------------------------------------- topology
// MPI_Dims_create as used by topology.jl:28: zero entries are filled with a balanced factorisation (factors as
// close to each other as possible) in non-increasing order; non-zero entries are kept.
extern "C" int chmy_dims_create(int nranks, int ndims, int32_t* dims) {
    CHMY_REQUIRE(dims && ndims >= 1 && ndims <= 3 && nranks >= 1, "bad argument");
    long long fixed = 1;
    int nfree = 0;
    for (int a = 0; a < ndims; ++a) {
        CHMY_REQUIRE(dims[a] >= 0, "negative dims entry");
        if (dims[a] > 0) fixed *= dims[a]; else ++nfree;
    }
    CHMY_REQUIRE(nranks % fixed == 0, "nranks %d is not divisible by the fixed dims", nranks);
    int rest = (int)(nranks / fixed);
    if (nfree == 0) { CHMY_REQUIRE(rest == 1, "prod(dims) != nranks"); return CHMY_OK; }
    int vals[3] = {1, 1, 1};
    int primes[32], np = 0;
    for (int p = 2; (long long)p * p <= rest; ++p)
        while (rest % p == 0) { primes[np++] = p; rest /= p; }
    if (rest > 1) primes[np++] = rest;
    for (int q = np - 1; q >= 0; --q) {          // largest prime first onto the currently smallest factor
        int m = 0;
        for (int a = 1; a < nfree; ++a) if (vals[a] < vals[m]) m = a;
        vals[m] *= primes[q];
    }
    for (int a = 0; a < nfree; ++a)              // sort non-increasing
        for (int b = a + 1; b < nfree; ++b)
            if (vals[b] > vals[a]) { int t = vals[a]; vals[a] = vals[b]; vals[b] = t; }
    int q = 0;
    for (int a = 0; a < ndims; ++a) if (dims[a] == 0) dims[a] = vals[q++];
    return CHMY_OK;
}

extern "C" int chmy_comm_unique_id(uint8_t out[CHMY_UNIQUE_ID_BYTES]) {
    CHMY_REQUIRE(out != nullptr, "out is NULL");
    CHMY_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == CHMY_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    CHMY_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return CHMY_OK;
}

// CartesianTopology(comm, dims): MPI.Cart_create (non-periodic, row-major ranks), Cart_coords, Cart_shift(dim, 1)
// (topology.jl:26-41).
extern "C" int chmy_topo_create(chmy_ctx* ctx, int nranks, int rank, int ndims, const int32_t* dims,
                                const uint8_t unique_id[CHMY_UNIQUE_ID_BYTES]) {
    CHMY_REQUIRE(ctx && dims, "NULL argument");
    CHMY_REQUIRE(ctx->comm == nullptr, "this architecture already has a topology");
    CHMY_REQUIRE(ndims >= 1 && ndims <= 3 && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
    long long prod = 1;
    for (int a = 0; a < ndims; ++a) { CHMY_REQUIRE(dims[a] >= 1, "dims must be positive (call chmy_dims_create first)"); prod *= dims[a]; }
    CHMY_REQUIRE(prod == nranks, "prod(dims) = %lld != nranks = %d", prod, nranks);
    chmy_comm* c = (chmy_comm*)calloc(1, sizeof(chmy_comm));
    if (!c) { chmy_set_error("out of host memory"); return CHMY_ERR_NOMEM; }
    c->nranks = nranks; c->rank = rank; c->nd = ndims;
    int r = rank;
    for (int a = ndims - 1; a >= 0; --a) { c->dims[a] = dims[a]; c->coords[a] = r % dims[a]; r /= dims[a]; }
    for (int a = 0; a < 3; ++a) {
        c->nb[a][0] = c->nb[a][1] = -1;
        if (a >= ndims) continue;
        for (int s = 0; s < 2; ++s) {
            int cc[3] = {c->coords[0], c->coords[1], c->coords[2]};
            cc[a] += s == 0 ? -1 : 1;
            if (cc[a] < 0 || cc[a] >= dims[a]) continue;
            int nr = 0;
            for (int b = 0; b < ndims; ++b) nr = nr * dims[b] + cc[b];
            c->nb[a][s] = nr;
        }
    }
    if (nranks > 1) {
        CHMY_REQUIRE(unique_id != nullptr, "unique_id is NULL");
        int rc = nccl_load();
        if (rc != CHMY_OK) { free(c); return rc; }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        CHMY_CUDA(cudaSetDevice(ctx->device));
        ncclResult_t nr = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
        if (nr != ncclSuccess) {
            chmy_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(nr));
            free(c);
            return CHMY_ERR_NCCL;
        }
    }
    CHMY_CUDA(cudaMalloc(&c->d_scal, 64 * sizeof(double)));
    for (int a = 0; a < 3; ++a)
        for (int s = 0; s < 2; ++s) { memset(&c->link[a][s], 0, sizeof(PlLink)); c->link[a][s].peer = c->nb[a][s]; }
    const char* xm = getenv("CHMY_EXCHANGE");
    c->xmode = (xm && (xm[0] == 'p' || xm[0] == 'P' || xm[0] == '1')) ? CHMY_EXCHANGE_PEER : CHMY_EXCHANGE_NCCL;
    const char* to = getenv("CHMY_PEER_TIMEOUT_S");
    const double to_s = to ? atof(to) : 20.0;
    c->timeout_ns = (unsigned long long)((to_s > 0.0 ? to_s : 20.0) * 1.0e9);
    ctx->comm = c;
    return CHMY_OK;
}

int chmy_comm_destroy(chmy_comm* c) {
    if (!c) return CHMY_OK;
    for (int s = 0; s < 2; ++s) { cudaFree(c->sbuf[s]); cudaFree(c->rbuf[s]); }
    cudaFree(c->d_scal);
    for (int a = 0; a < 3; ++a)
        for (int s = 0; s < 2; ++s) {
            if (c->link[a][s].remote) cudaIpcCloseMemHandle(c->link[a][s].remote);
            if (c->link[a][s].local) cudaFree(c->link[a][s].local);
        }
    for (int q = 0; q < c->ngrave; ++q) cudaFree(c->grave[q]);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->nccl) g_nccl.CommDestroy(c->nccl);
    free(c);
    return CHMY_OK;
}

extern "C" int chmy_topo_coords(const chmy_ctx* ctx, int32_t coords[CHMY_MAX_DIMS]) {
    CHMY_REQUIRE(ctx && coords, "NULL argument");
    if (!ctx->comm) { chmy_set_error("architecture has no topology"); return CHMY_ERR_STATE; }
    for (int a = 0; a < 3; ++a) coords[a] = a < ctx->comm->nd ? ctx->comm->coords[a] : 0;
    return CHMY_OK;
}

extern "C" int chmy_topo_neighbors(const chmy_ctx* ctx, int32_t nb[CHMY_MAX_DIMS][2]) {
    CHMY_REQUIRE(ctx && nb, "NULL argument");
    if (!ctx->comm) { chmy_set_error("architecture has no topology"); return CHMY_ERR_STATE; }
    for (int a = 0; a < 3; ++a) for (int s = 0; s < 2; ++s) nb[a][s] = ctx->comm->nb[a][s];
    return CHMY_OK;
}

// MPI.Allreduce(x, MPI.MAX, comm) on scalars (stokes_3d_inc_ve_T_mpi_perf.jl:16-19)
extern "C" int chmy_allreduce_max(chmy_ctx* ctx, double* inout, int n) {
    CHMY_REQUIRE(ctx && inout && n >= 0 && n <= 64, "bad argument");
    if (!ctx->comm || ctx->comm->nranks == 1 || n == 0) return CHMY_OK;
    chmy_comm* c = ctx->comm;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaMemcpyAsync(c->d_scal, inout, n * sizeof(double), cudaMemcpyHostToDevice, ctx->s_main));
    CHMY_NCCL(g_nccl.AllReduce(c->d_scal, c->d_scal, (size_t)n, ncclDouble, ncclMax, c->nccl, ctx->s_main));
    CHMY_CUDA(cudaMemcpyAsync(inout, c->d_scal, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}

extern "C" int chmy_barrier(chmy_ctx* ctx) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    CHMY_TRY(chmy_synchronize(ctx));
    double x = 0.0;
    return chmy_allreduce_max(ctx, &x, 1);
}

// ---------------------------------------------------------------------------------------------- exchange
static int ensure_bufs(chmy_comm* c, int s, size_t elems) {
    if (c->cap[s] >= elems) return CHMY_OK;
    CHMY_CUDA(cudaFree(c->sbuf[s]));
    CHMY_CUDA(cudaFree(c->rbuf[s]));
    c->sbuf[s] = c->rbuf[s] = nullptr;
    c->cap[s] = 0;
    CHMY_CUDA(cudaMalloc(&c->sbuf[s], elems * sizeof(double)));
    CHMY_CUDA(cudaMalloc(&c->rbuf[s], elems * sizeof(double)));
    c->cap[s] = elems;
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- peer-store exchange
// CHMY_EXCHANGE_PEER (opt-in; peer_link.cuh has the protocol): the pack kernel writes straight into the neighbour's HBM over
// NVLink and one sequence flag per direction replaces ncclSend / ncclRecv.  NCCL is still used once per link to swap the
// CUDA IPC handles.  A link whose block cannot be mapped on either end falls back to NCCL for good (both ends agree).
struct PlFlagArgs {
    uint64_t*          post[2];
    uint64_t           post_val[2];
    const uint64_t*    wait[2];
    uint64_t           wait_val[2];
    unsigned long long timeout_ns;
    int*               err;
};

__device__ __forceinline__ void pl_st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t pl_ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pl_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One thread: post the flags (after everything this stream has written, peer memory included), then wait for the local
// ones.  A wait that outlasts timeout_ns raises *err and returns: a lost neighbour must not hang the device.
__global__ void k_pl_flags(const PlFlagArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int s = 0; s < 2; ++s)
        if (a.post[s]) { __threadfence_system(); pl_st_release_sys(a.post[s], a.post_val[s]); }
    for (int s = 0; s < 2; ++s) {
        if (!a.wait[s]) continue;
        const unsigned long long t0 = pl_now_ns();
        while (pl_ld_acquire_sys(a.wait[s]) < a.wait_val[s]) {
            if (pl_now_ns() - t0 > a.timeout_ns) {
                *(volatile int*)a.err = 1 + s;
                __threadfence_system();
                return;
            }
            __nanosleep(100);
        }
    }
}

int chmy_comm_check(const chmy_comm* c) {
    if (c && c->h_err && *(volatile int*)c->h_err) {
        chmy_set_error("peer-store halo exchange: waiting for the neighbour on side %d timed out", *(volatile int*)c->h_err);
        return CHMY_ERR_STATE;
    }
    return CHMY_OK;
}

static int pl_resources(chmy_comm* c) {
    if (c->d_stage) return CHMY_OK;
    CHMY_CUDA(cudaMalloc(&c->d_stage, 256));
    CHMY_CUDA(cudaHostAlloc(&c->h_stage, 256, cudaHostAllocDefault));
    CHMY_CUDA(cudaHostAlloc(&c->h_err, sizeof(int), cudaHostAllocMapped));
    *c->h_err = 0;
    CHMY_CUDA(cudaHostGetDevicePointer(&c->d_err, c->h_err, 0));
    return CHMY_OK;
}

// swap up to 128 bytes with a neighbour (host to host, through the stream and NCCL); blocks until both have theirs
static int pl_swap(chmy_comm* c, int peer, const void* out, void* in, size_t n, cudaStream_t st) {
    CHMY_REQUIRE(n <= 128, "hand-shake message too long");
    memcpy(c->h_stage, out, n);
    CHMY_CUDA(cudaMemcpyAsync(c->d_stage, c->h_stage, n, cudaMemcpyHostToDevice, st));
    CHMY_NCCL(g_nccl.GroupStart());
    CHMY_NCCL(g_nccl.Recv(c->d_stage + 128, n, ncclChar, peer, c->nccl, st));
    CHMY_NCCL(g_nccl.Send(c->d_stage, n, ncclChar, peer, c->nccl, st));
    CHMY_NCCL(g_nccl.GroupEnd());
    CHMY_CUDA(cudaMemcpyAsync(c->h_stage + 128, c->d_stage + 128, n, cudaMemcpyDeviceToHost, st));
    CHMY_CUDA(cudaStreamSynchronize(st));
    memcpy(in, c->h_stage + 128, n);
    return CHMY_OK;
}

static void pl_bury(chmy_comm* c, void* blk) {
    if (!blk) return;
    if (c->ngrave < 64) c->grave[c->ngrave++] = blk;     // beyond that the block leaks until the process ends
}

struct PlHello {
    cudaIpcMemHandle_t h;
    uint64_t           cap;
    uint64_t           cookie;      // what the block holds at PL_OFF_COOKIE
    int32_t            ok;
    int32_t            pad;
};
static_assert(sizeof(PlHello) <= 128, "PlHello must fit the staging line");

// Makes link (D, s) ready for a message of need_bytes.  Both ends of a link see the same message sizes, so both get here
// -- first use, or growth -- in the same exchange and pair their hand-shakes.
static int pl_link_ensure(chmy_comm* c, int D, int s, size_t need_bytes, cudaStream_t st) {
    PlLink& l = c->link[D][s];
    if (l.mode == PL_MODE_NCCL) return CHMY_OK;
    if (l.mode == PL_MODE_PEER && need_bytes <= l.cap) return CHMY_OK;
    CHMY_TRY(pl_resources(c));
    const size_t cap = pl_grow_cap(l.mode == PL_MODE_PEER ? l.cap : 0, need_bytes);
    PlHello mine, theirs;
    memset(&mine, 0, sizeof(mine));
    memset(&theirs, 0, sizeof(theirs));
    char* blk = nullptr;
    cudaError_t e = cudaMalloc(&blk, pl_block_bytes(cap));
    if (e == cudaSuccess) e = cudaMemsetAsync(blk, 0, pl_block_bytes(cap), st);     // flags start at 0; ordered before the swap
    // the nonce: distinct per rank, link and (re)allocation
    mine.cookie = 0x9e3779b97f4a7c15ull * (uint64_t)(c->rank + 1) ^ ((uint64_t)(D * 2 + s + 1) << 56) ^ (uint64_t)(uintptr_t)blk ^ (uint64_t)cap;
    if (e == cudaSuccess) e = cudaMemcpyAsync(blk + PL_OFF_COOKIE, &mine.cookie, sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.h, blk);
    if (e != cudaSuccess) (void)cudaGetLastError();
    mine.ok = e == cudaSuccess;
    mine.cap = cap;
    CHMY_TRY(pl_swap(c, l.peer, &mine, &theirs, sizeof(PlHello), st));
    char* remote = nullptr;
    int32_t mapped = 0, peer_mapped = 0;
    if (mine.ok && theirs.ok && theirs.cap == cap) {
        e = cudaIpcOpenMemHandle((void**)&remote, theirs.h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { (void)cudaGetLastError(); remote = nullptr; }
        if (remote) {        // does the mapping address the block the peer meant (and can this device read it)?
            uint64_t got = 0;
            e = cudaMemcpy(&got, remote + PL_OFF_COOKIE, sizeof(got), cudaMemcpyDefault);
            if (e != cudaSuccess) (void)cudaGetLastError();
            if (e != cudaSuccess || got != theirs.cookie) { cudaIpcCloseMemHandle(remote); remote = nullptr; }
        }
        mapped = remote != nullptr;
    }
    CHMY_TRY(pl_swap(c, l.peer, &mapped, &peer_mapped, sizeof(mapped), st));
    // retire what the link had: my stores into the old mapping precede the synchronised swaps in stream order
    if (l.remote) cudaIpcCloseMemHandle(l.remote);
    pl_bury(c, l.local);
    l.local = l.remote = nullptr;
    l.cap = 0;
    l.seq = 0;
    if (mapped && peer_mapped) {
        l.local = blk; l.remote = remote; l.cap = cap; l.mode = PL_MODE_PEER;
    } else {
        if (remote) cudaIpcCloseMemHandle(remote);
        pl_bury(c, blk);
        l.mode = PL_MODE_NCCL;
    }
    return CHMY_OK;
}

// the three stream operations of pl_exchange_dim (peer_link.cuh) as kernel launches on `st`
struct CudaPlOps {
    chmy_ctx*                     ctx;
    chmy_comm*                    c;
    int                           D;
    const chmy_batch_desc* const* side;
    cudaStream_t                  st;

    int flags(PlFlagArgs& a) const {
        a.timeout_ns = c->timeout_ns;
        a.err        = c->d_err;
        k_pl_flags<<<1, 32, 0, st>>>(a);
        ctx->n_launches++;
        CHMY_CUDA(cudaGetLastError());
        return CHMY_OK;
    }
    int push(int s, PlLink& l, int slot) const {
        c->n_peer_msgs++;
        return chmy_pack_fields(ctx, D, s, side[s]->nfields, side[s]->fields, l.remote + pl_off_slot(slot, l.cap), st);
    }
    int post_and_wait(PlLink* const l[2], const uint64_t k[2]) const {
        PlFlagArgs a;
        memset(&a, 0, sizeof(a));
        for (int s = 0; s < 2; ++s)
            if (l[s]) {
                a.post[s] = pl_flag(l[s]->remote, PL_OFF_DATA); a.post_val[s] = k[s];
                a.wait[s] = pl_flag(l[s]->local, PL_OFF_DATA);  a.wait_val[s] = k[s];
            }
        return flags(a);
    }
    int unpack(int s, PlLink& l, int slot) const {
        return chmy_unpack_fields(ctx, D, s, side[s]->nfields, side[s]->fields, l.local + pl_off_slot(slot, l.cap), st);
    }
};

extern "C" int chmy_set_exchange_mode(chmy_ctx* ctx, int mode) {
    CHMY_REQUIRE(ctx != nullptr, "ctx is NULL");
    CHMY_REQUIRE(mode == CHMY_EXCHANGE_NCCL || mode == CHMY_EXCHANGE_PEER, "bad exchange mode %d", mode);
    if (!ctx->comm) { chmy_set_error("architecture has no topology"); return CHMY_ERR_STATE; }
    CHMY_TRY(chmy_flush(ctx));
    ctx->comm->xmode = mode;
    return CHMY_OK;
}

extern "C" int chmy_exchange_stats(const chmy_ctx* ctx, uint64_t* peer_msgs, uint64_t* nccl_msgs) {
    CHMY_REQUIRE(ctx && peer_msgs && nccl_msgs, "NULL argument");
    *peer_msgs = ctx->comm ? ctx->comm->n_peer_msgs : 0;
    *nccl_msgs = ctx->comm ? ctx->comm->n_nccl_msgs : 0;
    return CHMY_OK;
}

// Both sides of one dimension: pack -> {send, recv} x sides in one NCCL group -> unpack, all on `st`.
// Slab geometry: communication_views.jl:1-34; message pairing: my side S talks to the neighbour's side 1-S.
int chmy_exchange_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int D, const chmy_batch_desc* left,
                      const chmy_batch_desc* right, cudaStream_t st) {
    (void)g;
    const chmy_batch_desc* side[2] = {left, right};
    chmy_comm* c = ctx->comm;
    if (!c) { chmy_set_error("halo exchange requested but the architecture has no topology"); return CHMY_ERR_STATE; }
    CHMY_TRY(chmy_comm_check(c));
    size_t len[2] = {0, 0}, esz[2] = {8, 8};
    ncclDataType_t ty[2] = {ncclDouble, ncclDouble};      // slabs travel in the fields' element type
    for (int s = 0; s < 2; ++s) {
        if (!side[s] || side[s]->kind != CHMY_BATCH_EXCHANGE) continue;
        if (c->nb[D][s] < 0) { chmy_set_error("no neighbor to communicate (dim %d, side %d)", D + 1, s + 1); return CHMY_ERR_STATE; }  // exchange_halo.jl:19
        CHMY_REQUIRE(side[s]->nfields >= 1 && side[s]->nfields <= CHMY_MAX_BATCH_FIELDS, "ExchangeBatch with %d fields", side[s]->nfields);
        for (int q = 0; q < side[s]->nfields; ++q) {
            CHMY_REQUIRE(side[s]->fields[q] != nullptr, "ExchangeBatch: NULL field");
            len[s] += (size_t)chmy_slab_len(side[s]->fields[q], D);
        }
        if (side[s]->fields[0]->dtype == CHMY_F32) { ty[s] = ncclFloat; esz[s] = 4; }   // chmy_pack_fields checks that the batch is uniform
    }
    if (len[0] + len[1] == 0) return CHMY_OK;
    // links that carry this dimension's messages as peer stores (CHMY_EXCHANGE_PEER); the others go through NCCL
    PlLink* pl[2] = {nullptr, nullptr};
    if (c->xmode == CHMY_EXCHANGE_PEER)
        for (int s = 0; s < 2; ++s) {
            if (!len[s]) continue;
            CHMY_TRY(pl_link_ensure(c, D, s, len[s] * esz[s], st));
            if (c->link[D][s].mode == PL_MODE_PEER) pl[s] = &c->link[D][s];
        }
    bool any_nccl = false;
    for (int s = 0; s < 2; ++s) {
        if (!len[s] || pl[s]) continue;
        CHMY_TRY(ensure_bufs(c, s, len[s]));
        CHMY_TRY(chmy_pack_fields(ctx, D, s, side[s]->nfields, side[s]->fields, c->sbuf[s], st));
        any_nccl = true;
    }
    if (any_nccl) {
        CHMY_NCCL(g_nccl.GroupStart());
        for (int s = 0; s < 2; ++s) {
            if (!len[s] || pl[s]) continue;
            CHMY_NCCL(g_nccl.Recv(c->rbuf[s], len[s], ty[s], c->nb[D][s], c->nccl, st));
            CHMY_NCCL(g_nccl.Send(c->sbuf[s], len[s], ty[s], c->nb[D][s], c->nccl, st));
            c->n_nccl_msgs++;
        }
        CHMY_NCCL(g_nccl.GroupEnd());
    }
    if (pl[0] || pl[1]) {
        CudaPlOps ops{ctx, c, D, side, st};
        CHMY_TRY(pl_exchange_dim(ops, pl));
    }
    for (int s = 0; s < 2; ++s)
        if (len[s] && !pl[s]) CHMY_TRY(chmy_unpack_fields(ctx, D, s, side[s]->nfields, side[s]->fields, c->rbuf[s], st));
    return CHMY_OK;
}

static void fill_exchange_batch(chmy_batch_desc* b, int nf, chmy_field* const* fs) {
    memset(b, 0, sizeof(*b));
    b->kind = CHMY_BATCH_EXCHANGE;
    b->nfields = nf;
    for (int q = 0; q < nf; ++q) b->fields[q] = fs[q];
}

// exchange_halo!(side, dim, arch, grid, fields...)  -- exchange_halo.jl:13-61
extern "C" int chmy_exchange_halo(chmy_ctx* ctx, const chmy_grid_desc* g, int dim, int side, int nf,
                                  chmy_field* const* fields, int flags) {
    CHMY_REQUIRE(ctx && g && fields, "NULL argument");
    CHMY_REQUIRE(dim >= 0 && dim < g->ndims && (side == 0 || side == 1), "bad (dim, side)");
    CHMY_REQUIRE(nf >= 1 && nf <= CHMY_MAX_BATCH_FIELDS, "bad field count %d", nf);
    chmy_batch_desc b[2];
    memset(b, 0, sizeof(b));
    fill_exchange_batch(&b[side], nf, fields);
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_TRY(chmy_exchange_dim(ctx, g, dim, &b[0], &b[1], ctx->s_main));
    if (flags & CHMY_LAUNCH_BLOCKING) CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}

// exchange_halo!(arch, grid, fields...): D = N..1, the Connected sides only -- exchange_halo.jl:73-84
extern "C" int chmy_exchange_halo_all(chmy_ctx* ctx, const chmy_grid_desc* g, int nf, chmy_field* const* fields, int flags) {
    CHMY_REQUIRE(ctx && g && fields, "NULL argument");
    CHMY_REQUIRE(nf >= 1 && nf <= CHMY_MAX_BATCH_FIELDS, "bad field count %d", nf);
    CHMY_TRY(chmy_flush(ctx));
    CHMY_CUDA(cudaSetDevice(ctx->device));
    for (int D = g->ndims - 1; D >= 0; --D) {
        chmy_batch_desc b[2];
        memset(b, 0, sizeof(b));
        bool any = false;
        for (int s = 0; s < 2; ++s)
            if (g->connectivity[D][s] == CHMY_CONNECTED) { fill_exchange_batch(&b[s], nf, fields); any = true; }
        if (any) CHMY_TRY(chmy_exchange_dim(ctx, g, D, &b[0], &b[1], ctx->s_main));
    }
    if (flags & CHMY_LAUNCH_BLOCKING) CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    return CHMY_OK;
}
