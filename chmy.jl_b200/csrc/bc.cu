// bc.cu -- boundary-condition batches, halo slab pack/unpack and the small field utilities (sm_100a).
//
// Boundary rules: src/BoundaryConditions/first_order_boundary_condition.jl:34-84 on a uniform grid
//   Dirichlet, field Vertex along D : f[b] = v                          b  = 1 | d      (the boundary node)
//   Dirichlet, field Center along D : f[h] = muladd(2, v - f[nb], f[nb])  h = 0 | d+1 ; nb = 1 | d
//   Neumann  , any location         : f[h] = muladd(spacing_D, -/+q, f[nb])
// Face range: src/BoundaryConditions/batch.jl:159-184 -- transverse index I_t = J-1 in 0..n_t+2, fields of a batch
// applied in batch order.  Halo slabs: src/Distributed/communication_views.jl:1-34.
#include "common.cuh"
#include "bc_kernels.cuh"      // BcEntry, BcBatchDev, bc_point ; SlabEntry, SlabBatch, slab_point (shared with the host emulation)

template <class T>
static BckView<T> bck_view(const chmy_field* f) {
    return BckView<T>{reinterpret_cast<T*>(f->p0), f->nd > 1 ? f->stride[1] : 0, f->nd > 2 ? f->stride[2] : 0};
}

// ---------------------------------------------------------------------------------------------- BC batches
// One thread per face point; both sides and all fields of a dimension in one launch (bc_point).
template <class T>
__global__ void __launch_bounds__(256) k_bc_dim(const BcBatchDev<T> b) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;   // first transverse index  (0..nt0-1)
    const int c = blockIdx.y;                              // second transverse index (0..nt1-1)
    if (a >= b.nt[0]) return;
    bc_point(b, a, c);
}

// Everything a batch must satisfy, checked BEFORE the launch it belongs to starts (chmy_validate_launch / chmy_bc): a
// descriptor that is wrong in its batches must not leave a half-executed launch behind.  Needs no device.
int chmy_validate_batch(const chmy_grid_desc* g, int dim, const chmy_batch_desc* b) {
    if (!b || b->kind == CHMY_BATCH_EMPTY) return CHMY_OK;
    CHMY_REQUIRE(b->nfields >= 0 && b->nfields <= CHMY_MAX_BATCH_FIELDS, "batch with %d fields (max %d)", b->nfields, CHMY_MAX_BATCH_FIELDS);
    if (b->kind == CHMY_BATCH_EXCHANGE) CHMY_REQUIRE(b->nfields >= 1, "ExchangeBatch without fields");
    for (int q = 0; q < b->nfields; ++q) {
        const chmy_field* f = b->fields[q];
        CHMY_REQUIRE(f != nullptr && f->nd == g->ndims, "batch (dim %d): bad field %d", dim + 1, q);
        CHMY_REQUIRE(f->dtype == b->fields[0]->dtype, "batch (dim %d): the fields must share one element type", dim + 1);
        for (int a = 0; a < g->ndims; ++a)
            CHMY_REQUIRE(f->d[a] == g->n[a] + (f->loc[a] == CHMY_VERTEX ? 1 : 0), "batch (dim %d): field %d does not match the grid", dim + 1, q);
        if (b->kind != CHMY_BATCH_FIELD) continue;
        CHMY_REQUIRE(b->bc_kind[q] == CHMY_DIRICHLET || b->bc_kind[q] == CHMY_NEUMANN, "FieldBatch: bad bc kind");
        if (const chmy_field* vf = b->value_field[q]) {
            CHMY_REQUIRE(g->ndims >= 2 && vf->nd == g->ndims - 1, "Field-valued condition: the value field must have %d dims", g->ndims - 1);
            CHMY_REQUIRE(vf->dtype == f->dtype, "Field-valued condition: the value field must have the field's element type");
            int t = 0;
            for (int a = 0; a < g->ndims; ++a) {
                if (a == dim) continue;
                CHMY_REQUIRE(vf->d[t] >= g->n[a], "Field-valued condition: value field too small along transverse dim %d", t + 1);
                ++t;
            }
        }
    }
    return CHMY_OK;
}

template <class T>
static int run_bc_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int dim, const chmy_batch_desc* left,
                      const chmy_batch_desc* right, int dtype, cudaStream_t st) {
    BcBatchDev<T> b;
    memset(&b, 0, sizeof(b));
    b.dim     = dim;
    b.spacing = (T)g->spacing[dim];
    const chmy_batch_desc* sides[2] = {left, right};
    for (int s = 0; s < 2; ++s) {
        const chmy_batch_desc* bd = sides[s];
        if (!bd || bd->kind != CHMY_BATCH_FIELD) continue;
        CHMY_REQUIRE(bd->nfields >= 0 && bd->nfields <= CHMY_MAX_BATCH_FIELDS, "FieldBatch with %d fields (max %d)",
                     bd->nfields, CHMY_MAX_BATCH_FIELDS);
        for (int q = 0; q < bd->nfields; ++q) {
            const chmy_field* f = bd->fields[q];
            CHMY_REQUIRE(f != nullptr && f->alloc != nullptr && f->nd == g->ndims, "FieldBatch: bad field %d", q);
            CHMY_REQUIRE(f->dtype == dtype, "FieldBatch: the fields of one dimension's batches must share an element type");
            CHMY_REQUIRE(bd->bc_kind[q] == CHMY_DIRICHLET || bd->bc_kind[q] == CHMY_NEUMANN, "FieldBatch: bad bc kind");
            for (int a = 0; a < g->ndims; ++a)
                CHMY_REQUIRE(f->d[a] == g->n[a] + (f->loc[a] == CHMY_VERTEX ? 1 : 0), "FieldBatch: field/grid size mismatch");
            bd->fields[q]->frame_dirty(ctx->batch_sig);      // a halo of a Vertex field lies outside the ops' index range
            BcEntry<T>& e = b.e[b.n++];
            e.f = bck_view<T>(f); e.kind = bd->bc_kind[q]; e.vertex = f->loc[dim] == CHMY_VERTEX; e.d = (int)f->d[dim];
            e.side = s; e.value = (T)bd->value[q];
            e.vp = nullptr; e.vsy = 0;
            if (const chmy_field* vf = bd->value_field[q]) {
                CHMY_REQUIRE(g->ndims >= 2 && vf->nd == g->ndims - 1, "Field-valued condition: the value field must have %d dims", g->ndims - 1);
                int t = 0;
                for (int a = 0; a < g->ndims; ++a) {
                    if (a == dim) continue;
                    CHMY_REQUIRE(vf->d[t] >= g->n[a], "Field-valued condition: value field too small along transverse dim %d", t + 1);
                    ++t;
                }
                CHMY_REQUIRE(vf->dtype == dtype, "Field-valued condition: the value field must have the field's element type");
                e.vp = reinterpret_cast<const T*>(vf->p0); e.vsy = vf->nd > 1 ? vf->stride[1] : 0;
            }
        }
    }
    if (b.n == 0) return CHMY_OK;
    int t = 0;
    b.nt[0] = b.nt[1] = 1;
    for (int a = 0; a < g->ndims; ++a)
        if (a != dim) b.nt[t++] = (int)g->n[a] + 3;      // remove_dim(dim, nvertices + 2), batch.jl:181
    const dim3 blk(128, 1, 1);
    const dim3 grd((b.nt[0] + 127) / 128, b.nt[1], 1);
    k_bc_dim<T><<<grd, blk, 0, st>>>(b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

// ---- all dimensions, sides and fields of a batch set in ONE launch (bc_all_point, bc_kernels.cuh)
constexpr int BC_ALL_CY = 8;     // face points per thread along the slower face direction
template <class T>
__global__ void __launch_bounds__(128) k_bc_all(const BcAllDev<T> b) {
    const int z = b.act[blockIdx.z], s = z & 1, D = (z >> 1) % 3, q = z / 6;
    int nt[2] = {1, 1}, t = 0;
    for (int a = 0; a < b.nd; ++a)
        if (a != D) nt[t++] = b.n[a] + 3;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nt[0]) return;
    const int c0 = blockIdx.y * BC_ALL_CY, c1 = c0 + BC_ALL_CY < nt[1] ? c0 + BC_ALL_CY : nt[1];
#pragma unroll 4
    for (int c = c0; c < c1; ++c) bc_all_point(b, q, D, s, a, c);      // independent points: a few per thread instead of one
}

// Can the batch set run as one launch?  No exchange anywhere (a halo exchange sits between two dimensions), at most
// BCK_ALL_FIELDS distinct fields of one element type, no field twice in one (dim, side).  Fills the device table.
template <class T>
static bool make_bc_all(const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], int dtype, BcAllDev<T>* out,
                        chmy_field** touched, int* ntouched) {
    BcAllDev<T>& b = *out;
    memset(&b, 0, sizeof(b));
    b.nd = g->ndims;
    for (int a = 0; a < 3; ++a) { b.n[a] = a < g->ndims ? (int)g->n[a] : 0; b.spacing[a] = a < g->ndims ? (T)g->spacing[a] : (T)0; }
    for (int q = 0; q < BCK_ALL_FIELDS; ++q)
        for (int D = 0; D < 3; ++D)
            for (int s = 0; s < 2; ++s) b.fld[q].r[D][s].kind = -1;
    chmy_field* fl[BCK_ALL_FIELDS];
    for (int D = 0; D < g->ndims; ++D)
        for (int s = 0; s < 2; ++s) {
            const chmy_batch_desc& bd = bc[D][s];
            if (bd.kind == CHMY_BATCH_EXCHANGE) return false;
            if (bd.kind != CHMY_BATCH_FIELD) continue;
            for (int k = 0; k < bd.nfields; ++k) {
                chmy_field* f = bd.fields[k];
                if (!f || !f->alloc || f->dtype != dtype || f->nd != g->ndims) return false;
                int q = 0;
                while (q < b.nf && fl[q] != f) ++q;
                if (q == b.nf) {
                    if (b.nf == BCK_ALL_FIELDS) return false;
                    fl[b.nf++] = f;
                    b.fld[q].f = bck_view<T>(f);
                    for (int a = 0; a < 3; ++a) { b.fld[q].d[a] = (int)f->d[a]; b.fld[q].vertex[a] = a < f->nd && f->loc[a] == CHMY_VERTEX; }
                }
                BcRule<T>& r = b.fld[q].r[D][s];
                if (r.kind >= 0) return false;                      // the same field twice on one side: keep the sequential order
                r.kind = bd.bc_kind[k]; r.value = (T)bd.value[k]; r.vp = nullptr; r.vsy = 0;
                if (const chmy_field* vf = bd.value_field[k]) {
                    if (vf->dtype != dtype) return false;
                    r.vp = reinterpret_cast<const T*>(vf->p0); r.vsy = vf->nd > 1 ? vf->stride[1] : 0;
                }
            }
        }
    for (int q = 0; q < b.nf; ++q) touched[q] = fl[q];
    *ntouched = b.nf;
    for (int q = 0; q < b.nf; ++q)
        for (int D = 0; D < g->ndims; ++D)
            for (int s = 0; s < 2; ++s)
                if (b.fld[q].r[D][s].kind >= 0) b.act[b.nact++] = (unsigned char)((q * 3 + D) * 2 + s);
    return b.nact > 0;
}

template <class T>
static int run_bc_all_t(chmy_ctx* ctx, const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], int dtype, cudaStream_t st,
                        int* handled) {
    BcAllDev<T> b;
    chmy_field* touched[BCK_ALL_FIELDS];
    int nt = 0;
    if (!make_bc_all<T>(g, bc, dtype, &b, touched, &nt)) return CHMY_OK;
    for (int q = 0; q < nt; ++q) touched[q]->frame_dirty(ctx->batch_sig);
    int m0 = 1, m1 = 1;        // largest face extents over the dims
    for (int D = 0; D < g->ndims; ++D) {
        int e[2] = {1, 1}, t = 0;
        for (int a = 0; a < g->ndims; ++a)
            if (a != D) e[t++] = (int)g->n[a] + 3;
        m0 = e[0] > m0 ? e[0] : m0; m1 = e[1] > m1 ? e[1] : m1;
    }
    CHMY_REQUIRE((m1 + BC_ALL_CY - 1) / BC_ALL_CY <= 65535, "face too large for one launch");
    k_bc_all<T><<<dim3((m0 + 127) / 128, (m1 + BC_ALL_CY - 1) / BC_ALL_CY, b.nact), 128, 0, st>>>(b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    *handled = 1;
    return CHMY_OK;
}

// *handled = 1: the whole batch set ran as one launch; 0: the caller applies it dimension by dimension
int chmy_run_bc_all(chmy_ctx* ctx, const chmy_grid_desc* g, const chmy_batch_desc bc[CHMY_MAX_DIMS][2], cudaStream_t st, int* handled) {
    *handled = 0;
    int dtype = -1;
    for (int D = 0; D < g->ndims && dtype < 0; ++D)
        for (int s = 0; s < 2 && dtype < 0; ++s)
            if (bc[D][s].kind == CHMY_BATCH_FIELD && bc[D][s].nfields > 0 && bc[D][s].fields[0]) dtype = bc[D][s].fields[0]->dtype;
    if (dtype < 0) return CHMY_OK;
    if (dtype == CHMY_F32) return run_bc_all_t<float>(ctx, g, bc, dtype, st, handled);
    return run_bc_all_t<double>(ctx, g, bc, dtype, st, handled);
}

// one thread sleeping on a device counter: the fall-back of cuStreamWaitValue32 (api.cu, run_overlapped)
__global__ void k_spin_until(const unsigned int* counter, unsigned int target) {
    while (*reinterpret_cast<const volatile unsigned int*>(counter) < target) __nanosleep(500);
    __threadfence();
}
int chmy_spin_until(chmy_ctx* ctx, const unsigned int* counter, unsigned int target, cudaStream_t st) {
    k_spin_until<<<1, 1, 0, st>>>(counter, target);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

int chmy_run_bc_dim(chmy_ctx* ctx, const chmy_grid_desc* g, int dim, const chmy_batch_desc* left,
                    const chmy_batch_desc* right, cudaStream_t st) {
    int dtype = CHMY_F64;        // element type of the first field decides; run_bc_dim checks the rest against it
    const chmy_batch_desc* sides[2] = {left, right};
    for (int s = 1; s >= 0; --s)
        if (sides[s] && sides[s]->kind == CHMY_BATCH_FIELD && sides[s]->nfields > 0 && sides[s]->fields[0])
            dtype = sides[s]->fields[0]->dtype;
    if (dtype == CHMY_F32) return run_bc_dim<float>(ctx, g, dim, left, right, dtype, st);
    return run_bc_dim<double>(ctx, g, dim, left, right, dtype, st);
}

// ---------------------------------------------------------------------------------------------- halo slabs
// (geometry and per-element body: bc_kernels.cuh)
long long chmy_slab_len(const chmy_field* f, int dim) {
    long long len = 1;
    for (int a = 0; a < f->nd; ++a)
        if (a != dim) len *= f->sd[a];
    return len;
}

template <bool PACK, class T>
__global__ void __launch_bounds__(256) k_slab(const SlabBatch<T> b, T* __restrict__ buf) {
    slab_point<PACK, T>(b, buf, blockIdx.z, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y);
}

template <class T>
static int make_slab_batch(int dim, int side, int nf, chmy_field* const* fs, bool send, SlabBatch<T>* out) {
    CHMY_REQUIRE(nf >= 1 && nf <= CHMY_MAX_BATCH_FIELDS, "exchange with %d fields (max %d)", nf, CHMY_MAX_BATCH_FIELDS);
    SlabBatch<T>& b = *out;
    memset(&b, 0, sizeof(b));
    b.n = nf; b.dim = dim; b.nd = fs[0] ? fs[0]->nd : 0;
    long long off = 0;
    for (int q = 0; q < nf; ++q) {
        const chmy_field* f = fs[q];
        CHMY_REQUIRE(f != nullptr && f->alloc != nullptr && dim < f->nd, "exchange: bad field %d", q);
        const int ov = f->loc[dim] == CHMY_VERTEX ? 1 : 0;
        SlabEntry<T>& e = b.e[q];
        e.f   = bck_view<T>(f);
        e.idx = send ? (side == 0 ? 1 + ov : (int)f->d[dim] - ov) : (side == 0 ? 0 : (int)f->d[dim] + 1);
        int t = 0, ext[2] = {1, 1};
        for (int a = 0; a < f->nd; ++a)
            if (a != dim) ext[t++] = (int)f->sd[a];
        e.e0 = ext[0]; e.e1 = ext[1];
        e.off = off;
        off += chmy_slab_len(f, dim);
    }
    return CHMY_OK;
}

template <bool PACK, class T>
static int run_slab(chmy_ctx* ctx, int dim, int side, int nf, chmy_field* const* fs, T* dbuf, cudaStream_t st) {
    SlabBatch<T> b;
    for (int q = 0; q < nf; ++q)
        CHMY_REQUIRE(fs[q] != nullptr && fs[q]->nd == fs[0]->nd && fs[q]->dtype == fs[0]->dtype,
                     "the fields of one halo exchange must share dimensionality and element type");
    CHMY_TRY(make_slab_batch<T>(dim, side, nf, fs, PACK, &b));
    int m0 = 1, m1 = 1;
    for (int q = 0; q < nf; ++q) { m0 = b.e[q].e0 > m0 ? b.e[q].e0 : m0; m1 = b.e[q].e1 > m1 ? b.e[q].e1 : m1; }
    const dim3 blk(64, m1 > 1 ? 4 : 1, 1);
    const dim3 grd((m0 + blk.x - 1) / blk.x, (m1 + blk.y - 1) / blk.y, nf);
    k_slab<PACK, T><<<grd, blk, 0, st>>>(b, dbuf);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

int chmy_pack_fields(chmy_ctx* ctx, int dim, int side, int nf, chmy_field* const* fs, void* dbuf, cudaStream_t st) {
    CHMY_REQUIRE(nf >= 1 && fs && fs[0], "exchange: no fields");
    if (fs[0]->dtype == CHMY_F32) return run_slab<true, float>(ctx, dim, side, nf, fs, static_cast<float*>(dbuf), st);
    return run_slab<true, double>(ctx, dim, side, nf, fs, static_cast<double*>(dbuf), st);
}
int chmy_unpack_fields(chmy_ctx* ctx, int dim, int side, int nf, chmy_field* const* fs, const void* dbuf,
                       cudaStream_t st) {
    CHMY_REQUIRE(nf >= 1 && fs && fs[0], "exchange: no fields");
    for (int q = 0; q < nf; ++q)
        if (fs[q]) fs[q]->frame_dirty(ctx->batch_sig);
    if (fs[0]->dtype == CHMY_F32)
        return run_slab<false, float>(ctx, dim, side, nf, fs, static_cast<float*>(const_cast<void*>(dbuf)), st);
    return run_slab<false, double>(ctx, dim, side, nf, fs, static_cast<double*>(const_cast<void*>(dbuf)), st);
}

// ---------------------------------------------------------------------------------------------- field utilities
template <class T>
struct FillF {
    FVT<T> f; T v;
    __device__ void operator()(int i, int j, int k) const { fv_st(f, i, j, k, v); }
};
template <class T>
struct CopyF {
    FVT<T> d, s;
    __device__ void operator()(int i, int j, int k) const { fv_st(d, i, j, k, fv_ld(s, i, j, k)); }
};
template <class T>
struct InclF {
    FVT<T> f; InclDevT<T> q;
    __device__ void operator()(int i, int j, int k) const { fv_st(f, i, j, k, incl_eval(q, i, j, k)); }
};

// set!(C, grid, (x, y[, z]) -> exp(-x^2 - y^2 [- z^2])): the Gaussian initial condition of the diffusion drivers
// (examples/diffusion_2d_mpi.jl:46, diffusion_2d_mpi_perf.jl:56) evaluated where the field lives -- coordinates as in
// uniform_axis.jl:18-19 (muladd -> fma), -x^2 - y^2 parsed as (-(x*x)) - (y*y).  exp is CUDA's (<= 1 ulp): agrees with a
// host evaluation to ~2e-16 relative, not bit for bit.
template <class T>
struct GaussF {
    FVT<T> f; InclDevT<T> q;      // origin / spacing / loc of q are the grid's and the field's; the rest is unused
    __device__ void operator()(int i, int j, int k) const {
        const int I[3] = {i, j, k};
        T s = (T)0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d < q.nd) {
                const T c = coord_dev(q.origin[d], q.spacing[d], q.loc[d], I[d]);
                s = (d == 0) ? -(c * c) : s - c * c;
            }
        }
        fv_st(f, i, j, k, (T)exp(s));
    }
};

template <class F>
__global__ void __launch_bounds__(256) k_box_util(const F f, const Box b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i < b.n[0] && j < b.n[1]) f(b.lo[0] + i, b.lo[1] + j, b.lo[2] + k);
}

template <class F>
static int launch_util(chmy_ctx* ctx, const F& f, const Box& b, cudaStream_t st) {
    if (b.n[0] <= 0 || b.n[1] <= 0 || b.n[2] <= 0) return CHMY_OK;
    const dim3 blk(64, b.n[1] > 1 ? 4 : 1, 1);
    k_box_util<F><<<grid_for(b, blk), blk, 0, st>>>(f, b);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}

int chmy_box_from(const chmy_field* f, const int64_t* lo, const int64_t* hi, Box* out) {
    CHMY_REQUIRE(f->alloc != nullptr, "the field is descriptor-only (chmy_field_create_shell): it has no storage");
    for (int a = 0; a < 3; ++a) {
        if (a < f->nd) {
            CHMY_REQUIRE(lo[a] >= -1 && hi[a] <= f->d[a] + 2, "box [%lld,%lld] outside the padded field along dim %d",
                         (long long)lo[a], (long long)hi[a], a + 1);
            out->lo[a] = (int)lo[a];
            out->n[a]  = (int)(hi[a] - lo[a] + 1);
        } else {
            out->lo[a] = 0;
            out->n[a]  = 1;
        }
    }
    return CHMY_OK;
}

int chmy_fill_box(chmy_ctx* ctx, chmy_field* f, double v, const Box& b, cudaStream_t st) {
    f->frame_dirty(0);
    if (f->dtype == CHMY_F32) return launch_util(ctx, FillF<float>{f->viewT<float>(), (float)v}, b, st);
    return launch_util(ctx, FillF<double>{f->view(), v}, b, st);
}
int chmy_copy_box(chmy_ctx* ctx, chmy_field* d, const chmy_field* s, const Box& b, cudaStream_t st) {
    d->frame_dirty(0);
    if (d->dtype == CHMY_F32) return launch_util(ctx, CopyF<float>{d->viewT<float>(), s->viewT<float>()}, b, st);
    return launch_util(ctx, CopyF<double>{d->view(), s->view()}, b, st);
}
int chmy_incl_box(chmy_ctx* ctx, chmy_field* f, const InclDev& q, const Box& b, cudaStream_t st) {
    f->frame_dirty(0);
    return launch_util(ctx, InclF<double>{f->view(), q}, b, st);
}
int chmy_incl_box_f32(chmy_ctx* ctx, chmy_field* f, const InclDevT<float>& q, const Box& b, cudaStream_t st) {
    f->frame_dirty(0);
    return launch_util(ctx, InclF<float>{f->viewT<float>(), q}, b, st);
}
int chmy_gauss_box(chmy_ctx* ctx, chmy_field* f, const InclDev& q, const Box& b, cudaStream_t st) {
    f->frame_dirty(0);
    return launch_util(ctx, GaussF<double>{f->view(), q}, b, st);
}
int chmy_gauss_box_f32(chmy_ctx* ctx, chmy_field* f, const InclDevT<float>& q, const Box& b, cudaStream_t st) {
    f->frame_dirty(0);
    return launch_util(ctx, GaussF<float>{f->viewT<float>(), q}, b, st);
}

// ---------------------------------------------------------------------------------------------- max |f|
// maximum(abs.(interior(f))): exact and order-independent.  |x| is compared through its bit pattern as an unsigned
// integer (monotone for non-negative values of either element type; NaN patterns compare above +Inf, so a NaN propagates
// as in Julia).  Float32 fields reduce their 32-bit patterns; chmy_field_maxabs converts the winner back.
__device__ __forceinline__ unsigned long long abs_bits(double x) { return (unsigned long long)__double_as_longlong(fabs(x)); }
__device__ __forceinline__ unsigned long long abs_bits(float x) { return (unsigned long long)__float_as_uint(fabsf(x)); }

template <class T>
__global__ void __launch_bounds__(256) k_maxabs(const FVT<T> f, const Box b, unsigned long long* __restrict__ out) {
    unsigned long long m = 0ull;
    const long long rows   = (long long)b.n[1] * b.n[2];
    const int       lane   = threadIdx.x & 31;
    const int       warp   = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int       nwarps = (gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        const int j = b.lo[1] + (int)(r % b.n[1]);
        const int k = b.lo[2] + (int)(r / b.n[1]);
        for (int i = lane; i < b.n[0]; i += 32) {
            const unsigned long long v = abs_bits(fv_ld(f, b.lo[0] + i, j, k));
            m = v > m ? v : m;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    __shared__ unsigned long long sm[8];
    if (lane == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = sm[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long v = __shfl_xor_sync(0xffu, m, o);
            m = v > m ? v : m;
        }
        if (threadIdx.x == 0) atomicMax(out, m);
    }
}

int chmy_maxabs_box(chmy_ctx* ctx, const chmy_field* f, const Box& b, unsigned long long* d_out, cudaStream_t st) {
    const long long rows = (long long)b.n[1] * b.n[2];
    long long want = (rows + 7) / 8;                        // 8 warps per block, one row per warp per pass
    const long long cap = (long long)ctx->sm_count * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (f->dtype == CHMY_F32) k_maxabs<float><<<(unsigned)want, 256, 0, st>>>(f->viewT<float>(), b, d_out);
    else k_maxabs<double><<<(unsigned)want, 256, 0, st>>>(f->view(), b, d_out);
    ctx->n_launches++;
    CHMY_CUDA(cudaGetLastError());
    return CHMY_OK;
}
