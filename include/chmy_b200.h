/*
 * chmy_b200.h -- C ABI of the B200-native hot path of PTsolvers/Chmy.jl (v0.1.25).
 *
 * The reference has no FFI: its backend seam is Julia multiple dispatch on the KernelAbstractions backend
 * (ext/ChmyCUDAExt/ChmyCUDAExt.jl:1-25).  Because a C library cannot JIT arbitrary `@kernel` bodies, the seam moves
 * up to the callers of KernelAbstractions: each entry point below replaces one reference interface (cited
 * file:line, relative to the reference root) and is what a `ChmyB200Ext` Julia package `ccall`s.
 *
 * Conventions
 *   - every function returns 0 on success or a negative chmy_status; it never throws or aborts.
 *     chmy_last_error() gives a thread-local message (the Julia glue turns non-zero into `error(msg)`,
 *     matching the reference's plain `error`/`@assert`: exchange_halo.jl:19, stack_allocator.jl:42).
 *   - one chmy_ctx per GPU / rank process, used by one host thread at a time.
 *   - logical indices are the reference's: 1-based, I in 1..d is the interior, 0 and d+1 the halo, -1 and d+2
 *     zero padding (src/Fields/field.jl:18-22,56-62).  Boxes are inclusive [lo, hi] in logical indices.
 *   - the solver ops (CHMY_OP_COMPUTE_Q .. CHMY_OP_UPDATE_THERMAL) are Float64 programs, like the reference's example
 *     drivers.  Fields, set!/copies, maxabs, bc!, halo exchange and the grid operators (CHMY_OP_OPERATOR) exist for
 *     Float32 as well (chmy_field_create_typed), as the reference's tests instantiate them (test/common.jl:9).  Host
 *     buffers of the copy / halo entry points hold elements of the field's type; scalar arguments stay `double` (a
 *     Float32 value converts exactly) and are rounded to the field's type where the reference's would have it.  A
 *     condition value is therefore taken in eltype(field), as the reference's tests pass it (`Dirichlet(T(2.0))`,
 *     test/test_boundary_conditions.jl:34); a Float64 value given for a Float32 field is rounded first, where Julia would
 *     promote that one muladd to Float64 and round its result.
 *   - all work is stream-ordered on the context's streams; results are visible to the host after a blocking
 *     launch (CHMY_LAUNCH_BLOCKING, the reference's semantics, KernelLaunch.jl:117), any copy_to_host /
 *     maxabs call, or chmy_synchronize().
 */
#ifndef CHMY_B200_H
#define CHMY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHMY_ABI_VERSION 3

#define CHMY_MAX_DIMS 3
#define CHMY_MAX_BATCH_FIELDS 8
#define CHMY_MAX_OP_FIELDS 24
#define CHMY_MAX_SCALARS 8
#define CHMY_UNIQUE_ID_BYTES 128

typedef struct chmy_ctx chmy_ctx;       /* Architecture + its streams (+ topology and communicator) */
typedef struct chmy_field chmy_field;   /* Field{Float64,N,L,1}: padded device storage               */

typedef enum {
    CHMY_OK = 0,
    CHMY_ERR_ARG = -1,      /* invalid argument / descriptor                     */
    CHMY_ERR_CUDA = -2,     /* CUDA runtime error                                */
    CHMY_ERR_NCCL = -3,     /* NCCL error or NCCL not loadable                   */
    CHMY_ERR_STATE = -4,    /* e.g. exchange requested without a topology        */
    CHMY_ERR_NOMEM = -5
} chmy_status;

typedef enum { CHMY_CENTER = 0, CHMY_VERTEX = 1 } chmy_loc;            /* src/Grids/Grids.jl:26-43  */
typedef enum { CHMY_BOUNDED = 0, CHMY_CONNECTED = 1 } chmy_conn;       /* src/Grids/Grids.jl:52-57  */
typedef enum { CHMY_DIRICHLET = 0, CHMY_NEUMANN = 1 } chmy_bc_kind;    /* first_order_boundary_condition.jl:9-29 */
typedef enum { CHMY_BATCH_EMPTY = 0, CHMY_BATCH_FIELD = 1, CHMY_BATCH_EXCHANGE = 2 } chmy_batch_kind;  /* batch.jl:36-66 */

/* Field storage layout.  PITCHED (default): rows padded so that logical index 0 of every row is 128-byte
 * aligned (vector loads, TMA-legal strides).  DENSE: exactly the reference's dense column-major
 * array of dims+4 (field.jl:58-59).  Logical contents are identical; kernels accept both. */
typedef enum { CHMY_LAYOUT_PITCHED = 0, CHMY_LAYOUT_DENSE = 1 } chmy_layout;
typedef enum { CHMY_F64 = 0, CHMY_F32 = 1 } chmy_dtype;                /* eltype(field): TEST_TYPES, test/common.jl:9 */

/* The ops `launch` can run: the @kernel functions of the named example solvers. */
typedef enum {
    CHMY_OP_NONE = 0,            /* descriptor carries only boundary batches (bc!)                       */
    CHMY_OP_COMPUTE_Q = 1,       /* examples/diffusion_2d.jl:8-13     fields: q.x q.y C          scalars: chi      */
    CHMY_OP_UPDATE_C = 2,        /* examples/diffusion_2d.jl:15-19    fields: C q.x q.y          scalars: dt       */
    CHMY_OP_UPDATE_OLD = 3,      /* stokes_3d_inc_ve_T.jl:11-21       fields: T tau[nt] T_old tau_old[nt]          */
    CHMY_OP_UPDATE_STRESS = 4,   /* stokes_3d_inc_ve_T.jl:23-46       fields: tau[nt] Pr divV V[nd] tau_old[nt]
                                                                       scalars: eta eta_ve G dt dtau_Pr dtau_r      */
    CHMY_OP_UPDATE_VELOCITY = 5, /* stokes_3d_inc_ve_T.jl:48-57       fields: V[nd] r_V[nd] Pr tau[nt] rho_g|NULL
                                                                       scalars: eta_ve nudtau                       */
    CHMY_OP_UPDATE_THERMAL_FLUX = 6, /* stokes_3d_inc_ve_T.jl:59-71   fields: qT[nd] T V[nd]     scalars: lambda   */
    CHMY_OP_UPDATE_THERMAL = 7,  /* stokes_3d_inc_ve_T.jl:73-77       fields: T T_old qT[nd]     scalars: dt       */
    CHMY_OP_OPERATOR = 8         /* dst[I] = OPERATOR(src...)[I]: chmy_launch_desc::oper / oper_dim, see chmy_operator */
} chmy_op;                       /* nd = grid.ndims; nt = 3 (2D: xx yy xy) or 6 (3D: xx yy zz xy xz yz)            */

/* The staggered-grid operators of src/GridOperators as field-level kernels (CHMY_OP_OPERATOR, ABI v3).  The reference
 * exposes them as point functions called from user `@kernel`s (test/test_grid_operators.jl:21-126,
 * test/test_interpolations.jl:22-70); a C library cannot JIT those kernels, so each operator is offered as
 * `dst[I] = OP(src...)[I]` over the usual launch range I in [0, n+1]^N.  `dim` = chmy_launch_desc::oper_dim (0-based).
 * Locations are checked: the result of an operator lives where the reference's operator puts it.               */
typedef enum {
    CHMY_OPER_LEFT = 1,       /* left(f, dim, I)   GridOperators.jl:23-33, field_operators.jl:2-6    fields: dst f   dst at flipped(loc(f), dim) */
    CHMY_OPER_RIGHT = 2,      /* right(f, dim, I)  field_operators.jl:8-12                            fields: dst f   "                          */
    CHMY_OPER_DELTA = 3,      /* δ(f, dim, I)      partial_derivatives.jl:2                           fields: dst f   "                          */
    CHMY_OPER_PARTIAL = 4,    /* ∂(f, grid, dim, I) partial_derivatives.jl:5                          fields: dst f   "                          */
    CHMY_OPER_PARTIAL2 = 5,   /* ∂²(f, grid, dim, I) partial_derivatives.jl:7-12                      fields: dst f   dst at loc(f)              */
    CHMY_OPER_DKD = 6,        /* ∂k∂(f, k, grid, dim, I) partial_derivatives.jl:14-21                 fields: dst f k dst at loc(f), k anywhere  */
    CHMY_OPER_LERP = 7,       /* lerp(f, location(dst), grid, I) interpolation.jl:14,63-87            fields: dst f   any two locations          */
    CHMY_OPER_HLERP = 8,      /* hlerp(f, location(dst), grid, I) interpolation.jl:15,94              fields: dst f   "                          */
    CHMY_OPER_DIVG = 9,       /* divg(V, grid, I)  field_operators.jl:50-55                           fields: dst V[nd]  dst at flipped(loc(V.d), d) for every d */
    CHMY_OPER_LAPL = 10,      /* lapl(f, grid, I)  field_operators.jl:72-77                           fields: dst f   dst at loc(f)              */
    CHMY_OPER_DIVG_GRAD = 11, /* divg_grad(f, k, grid, I) field_operators.jl:95-100                   fields: dst f k dst at loc(f)              */
    CHMY_OPER_VMAG = 12,      /* vmag(V, grid, I)  field_operators.jl:116-121                         fields: dst V[nd]  dst at Center           */
    CHMY_OPER_GRAD = 13,      /* V.d[I] = ∂_d(f)   test_grid_operators.jl:24-30 (divg1!)              fields: V[nd] f  V.d at flipped(loc(f), d) */
    CHMY_OPER_KGRAD = 14      /* V.d[I] = lerp(k, location(V.d)) * ∂_d(f) test_grid_operators.jl:76-82 fields: V[nd] f k                         */
} chmy_operator;

/* UniformGrid: the numbers src/Grids/structured_grid.jl:27-39 (+ distributed_grid.jl:19-36) produces. */
typedef struct {
    int32_t ndims;
    int32_t _pad;
    int64_t n[CHMY_MAX_DIMS];               /* local number of cells                       */
    double  origin[CHMY_MAX_DIMS];
    double  extent[CHMY_MAX_DIMS];
    double  spacing[CHMY_MAX_DIMS];         /* uniform_axis.jl:8                           */
    double  inv_spacing[CHMY_MAX_DIMS];     /* uniform_axis.jl:9                           */
    int32_t connectivity[CHMY_MAX_DIMS][2]; /* chmy_conn per (dim, side)                   */
} chmy_grid_desc;

/* One side of one dimension of a BatchSet (src/BoundaryConditions/batch.jl:8,36-66). */
typedef struct {
    int32_t     kind;                                 /* chmy_batch_kind                                  */
    int32_t     nfields;
    chmy_field* fields[CHMY_MAX_BATCH_FIELDS];
    int32_t     bc_kind[CHMY_MAX_BATCH_FIELDS];       /* chmy_bc_kind (FIELD batches)                     */
    double      value[CHMY_MAX_BATCH_FIELDS];         /* `nothing` -> 0.0, Number -> value                */
    chmy_field* value_field[CHMY_MAX_BATCH_FIELDS];   /* Field-valued condition (first_order_boundary_condition.jl:38-40):
                                                         an (N-1)-dimensional Field read at remove_dim(dim, I); NULL ->
                                                         value[].  A BoundaryFunction (boundary_function.jl:28-44) is a
                                                         host closure: the binding evaluates it into such a Field. */
} chmy_batch_desc;

/* FunctionField with the drivers' `init_incl` body (function_field.jl:12-59,
 * stokes_3d_inc_ve_T_mpi_perf.jl:141-142): value = (sum_d (x_d - c0_d)^2 < r^2) ? in : out. */
typedef struct {
    int32_t active;
    int32_t loc[CHMY_MAX_DIMS];
    double  c0[CHMY_MAX_DIMS];
    double  r, in, out;
} chmy_inclusion;

typedef enum {
    CHMY_LAUNCH_ASYNC = 0,        /* stream-ordered; host returns immediately                              */
    CHMY_LAUNCH_BLOCKING = 1,     /* reference semantics: returns after completion (KernelLaunch.jl:117)   */
    CHMY_LAUNCH_EXACT_SPLIT = 2   /* honour outer_width literally instead of treating it as a hint         */
} chmy_launch_flags;

/* `launcher(arch, grid, op => args; bc)`: src/KernelLaunch.jl:105-119,152-183. */
typedef struct {
    int32_t         op;                                     /* chmy_op                                    */
    int32_t         flags;                                  /* chmy_launch_flags, OR-ed                   */
    chmy_grid_desc  grid;
    int32_t         nfields;
    int32_t         nscalars;
    chmy_field*     fields[CHMY_MAX_OP_FIELDS];
    double          scalars[CHMY_MAX_SCALARS];
    chmy_inclusion  rho_g;                                  /* UPDATE_VELOCITY when rho_g is a FunctionField */
    int32_t         has_bc;                                 /* bc === nothing ? 0 : 1                     */
    int32_t         has_outer_width;                        /* Launcher built with outer_width            */
    int64_t         outer_width[CHMY_MAX_DIMS];
    chmy_batch_desc bc[CHMY_MAX_DIMS][2];
    int32_t         oper;                                   /* chmy_operator (CHMY_OP_OPERATOR only; ABI v3) */
    int32_t         oper_dim;                               /* 0-based dim of LEFT..DKD                      */
} chmy_launch_desc;

typedef struct {
    int32_t ndims;
    int32_t layout;
    int32_t loc[CHMY_MAX_DIMS];
    int64_t dims[CHMY_MAX_DIMS];       /* logical size(grid, loc)                                         */
    int64_t stride[CHMY_MAX_DIMS];     /* element strides; stride[0] == 1                                 */
    void*   origin_ptr;                /* device address of logical index (1,1,1)                          */
    void*   base_ptr;                  /* device address of storage element (-1,-1,-1)                     */
    size_t  bytes;                     /* allocation size                                                  */
    int32_t dtype;                     /* chmy_dtype; strides count elements of this type (ABI v3)         */
    int32_t _pad;
} chmy_field_info;

/* ---- library ------------------------------------------------------------------------------------------ */
int         chmy_abi_version(void);
const char* chmy_last_error(void);
int         chmy_device_count(int* count);
size_t      chmy_struct_size(int which);   /* 0 grid_desc, 1 batch_desc, 2 inclusion, 3 launch_desc, 4 field_info:
                                              lets a binding verify its struct layout against the library        */

/* ---- Architecture: Arch(backend; device_id) src/Architectures.jl:46-49 ; activate! :71-74 ;
 *      ext/ChmyCUDAExt/ChmyCUDAExt.jl:15-17 (set_device!/get_device) ------------------------------------ */
int chmy_ctx_create(int device_id /* 1-based as in the reference */, chmy_ctx** out);
int chmy_ctx_destroy(chmy_ctx* ctx);
int chmy_ctx_device(const chmy_ctx* ctx, int* device_id);
int chmy_synchronize(chmy_ctx* ctx);                      /* KernelAbstractions.synchronize(backend)        */
int chmy_ctx_launch_count(const chmy_ctx* ctx, uint64_t* kernels); /* kernels launched so far (bench)       */
/* device-side timing (CUDA events recorded on the context's main stream): slot in 0..CHMY_MAX_EVENTS-1 */
#define CHMY_MAX_EVENTS 4096
int chmy_event_record(chmy_ctx* ctx, int slot);
int chmy_event_elapsed_ms(chmy_ctx* ctx, int slot_start, int slot_stop, float* ms);   /* synchronises on slot_stop */
/* Measurement aid: record event slot_begin right before and slot_end right after the kernel of every fused 3D sweep, on the
 * stream it is launched on (a launch with boundary batches is one API call but two kernels: this isolates the sweep).
 * (-1, -1) switches it off.  bench.py uses it for roofline.kernel_ms. */
int chmy_time_fused_sweep(chmy_ctx* ctx, int slot_begin, int slot_end);
/* raw handles for tools that time or capture the context's work (cudaStream_t as void*) */
int chmy_ctx_streams(const chmy_ctx* ctx, void** main_stream, void** boundary_stream);

/* ---- Distributed: CartesianTopology src/Distributed/topology.jl:26-41 ;
 *      Arch(backend, comm, dims) distributed_architecture.jl:27-34 -------------------------------------- */
int chmy_dims_create(int nranks, int ndims, int32_t* dims /* in: 0 = free, out: filled */); /* MPI.Dims_create */
int chmy_comm_unique_id(uint8_t out[CHMY_UNIQUE_ID_BYTES]);     /* rank 0; broadcast out-of-band (MPI.bcast)  */
int chmy_topo_create(chmy_ctx* ctx, int nranks, int rank, int ndims, const int32_t* dims,
                     const uint8_t unique_id[CHMY_UNIQUE_ID_BYTES]);
int chmy_topo_coords(const chmy_ctx* ctx, int32_t coords[CHMY_MAX_DIMS]);               /* cart_coords :90   */
int chmy_topo_neighbors(const chmy_ctx* ctx, int32_t nb[CHMY_MAX_DIMS][2]);             /* neighbors :99 ; -1 = PROC_NULL */
int chmy_allreduce_max(chmy_ctx* ctx, double* inout, int n);    /* MPI.Allreduce(x, MAX) in the drivers       */
int chmy_barrier(chmy_ctx* ctx);                                /* MPI.Barrier(cart_comm)                     */

/* ---- Fields: Field(backend, grid, loc; halo=1) src/Fields/field.jl:56-62 ------------------------------ */
int chmy_field_create(chmy_ctx* ctx, int ndims, const int64_t* dims, const int32_t* loc, int layout,
                      chmy_field** out);                              /* zero-initialised                  */
int chmy_field_create_typed(chmy_ctx* ctx, int ndims, const int64_t* dims, const int32_t* loc, int layout, int dtype,
                            chmy_field** out);                        /* Field(backend, grid, loc, T) field.jl:56 */
/* Descriptor-only field (no context, no device storage): accepted by chmy_validate_launch and chmy_field_get_info /
 * chmy_halo_slab_len, refused by everything that would touch storage.  Lets a binding check its flattening of
 * `op => args` against the library's own rules on a machine without a GPU. */
int chmy_field_create_shell(int ndims, const int64_t* dims, const int32_t* loc, int layout, int dtype, chmy_field** out);
int chmy_field_destroy(chmy_field* f);
int chmy_field_get_info(const chmy_field* f, chmy_field_info* out);
int chmy_field_fill(chmy_ctx* ctx, chmy_field* f, double value, const int64_t* lo, const int64_t* hi);
                       /* fill!(parent(f),v): lo=-1,hi=d+2 ; set!(f,v) field.jl:87: lo=1,hi=d               */
int chmy_field_copy_from_host(chmy_ctx* ctx, chmy_field* f, const void* src, const int64_t* lo, const int64_t* hi);
                       /* set!(f, A) field.jl:98: dense column-major host box of eltype(f)                  */
int chmy_field_copy_to_host(chmy_ctx* ctx, const chmy_field* f, void* dst, const int64_t* lo, const int64_t* hi);
                       /* Array(interior(f; with_halo)) field.jl:33-37                                      */
int chmy_field_copy(chmy_ctx* ctx, chmy_field* dst, const chmy_field* src, const int64_t* lo, const int64_t* hi);
                       /* set!(f, other) field.jl:109-119                                                   */
int chmy_field_set_inclusion(chmy_ctx* ctx, chmy_field* f, const chmy_grid_desc* grid, const chmy_inclusion* inc);
/* set!(f, grid, (x, y[, z]) -> exp(-x^2 - y^2 [- z^2])) (field.jl:131-142 with the Gaussian initial condition of
 * examples/diffusion_2d_mpi.jl:46): evaluated on the device at the field's coordinates; agrees with a host evaluation to
 * ~1e-16 relative (CUDA's exp), not bit for bit. */
int chmy_field_set_gaussian(chmy_ctx* ctx, chmy_field* f, const chmy_grid_desc* grid);
                       /* set!(f, grid, init_incl; parameters) field.jl:121-142 (interior only)             */
int chmy_field_maxabs(chmy_ctx* ctx, const chmy_field* f, const int64_t* lo, const int64_t* hi, double* out);
                       /* maximum(abs.(interior(f))) in the drivers, e.g. stokes_3d_inc_ve_T.jl:158,172-175 */
int chmy_field_maxabs_many(chmy_ctx* ctx, int n, const chmy_field* const* fields, const int64_t* lo, const int64_t* hi,
                           double* out);
                       /* the residual check's four maxima (stokes_3d_inc_ve_T.jl:171-175) in ONE round trip: n <= 64
                          reductions back to back, one copy, one synchronisation; lo / hi are n x CHMY_MAX_DIMS        */
/* Page-locked host memory for the host side of set!(f, A) / Array(interior(f)) (field.jl:98, :33-37): copies from/to
 * such a buffer run at full PCIe/C2C rate without a staging pass.  Any host pointer is accepted by the copy entry
 * points; these two only provide the fast kind (the Julia glue wraps it with unsafe_wrap(Array, ...)). */
int chmy_host_alloc(chmy_ctx* ctx, size_t bytes, void** out);
int chmy_host_free(chmy_ctx* ctx, void* p);

/* ---- the hot entry points ----------------------------------------------------------------------------- */
int chmy_launch(chmy_ctx* ctx, const chmy_launch_desc* desc);           /* src/KernelLaunch.jl:105-119       */
int chmy_validate_launch(const chmy_launch_desc* desc);                 /* chmy_launch's argument checks only: op id, field
                                                                           count / order / staggered locations / sizes /
                                                                           element types, batches; needs no device */
int chmy_bc(chmy_ctx* ctx, const chmy_grid_desc* grid,                  /* bc!(arch, grid, batchset)         */
            const chmy_batch_desc bc[CHMY_MAX_DIMS][2], int flags);     /*   src/BoundaryConditions/batch.jl:20-29 */
int chmy_exchange_halo(chmy_ctx* ctx, const chmy_grid_desc* grid, int dim, int side,   /* exchange_halo.jl:13-61 */
                       int nfields, chmy_field* const* fields, int flags);
int chmy_exchange_halo_all(chmy_ctx* ctx, const chmy_grid_desc* grid,                  /* exchange_halo.jl:73-84 */
                           int nfields, chmy_field* const* fields, int flags);

/* ---- transport of the halo exchange (exchange_halo.jl:13-61: Irecv + pack + Isend per field, host poll, unpack) ---------
 * CHMY_EXCHANGE_NCCL (default; measured on B200, profiles/): one pack kernel per (dim, side) for all fields, both sides of a
 *   dimension in one ncclSend/ncclRecv group, one unpack kernel per side.
 * CHMY_EXCHANGE_PEER (EXPERIMENTAL: protocol proven on the CPU by tests/test_peer_protocol.py, not yet run on a GPU): the
 *   pack kernel stores the slabs straight into a receive slot in the neighbour's HBM (mapped with CUDA IPC, written over
 *   NVLink) and one sequence flag per direction replaces the NCCL group; NCCL only swaps the IPC handles once per link.
 *   A link whose memory cannot be mapped on either end keeps using NCCL (both ends agree).  Waits give up after
 *   CHMY_PEER_TIMEOUT_S (default 20) and the next call on the context reports it.
 * Every rank of a topology must choose the same mode; results are identical.  Env: CHMY_EXCHANGE=nccl|peer. */
typedef enum { CHMY_EXCHANGE_NCCL = 0, CHMY_EXCHANGE_PEER = 1 } chmy_exchange_mode;
int chmy_set_exchange_mode(chmy_ctx* ctx, int mode);
int chmy_exchange_stats(const chmy_ctx* ctx, uint64_t* peer_msgs, uint64_t* nccl_msgs);   /* messages sent so far per transport */

/* ---- self-test and tuning hooks (used by tests/ and bench.py; not part of the reference's surface) -------- */
/* counts operands x (n pseudo-random ones from `seed`) for which the exact-division sequence used by the tuned
 * kernels for kernel-uniform divisors differs bitwise from IEEE x / c; *markstein_used = 0 when c is routed to the
 * true-division instantiation instead. */
int chmy_selftest_division(chmy_ctx* ctx, double c, long long n, unsigned long long seed,
                           unsigned long long* mismatches, int* markstein_used);
/* Division by a launch-uniform scalar inside the fused 3D sweep: is the two-operation sequence fma(x, RN(1/c), RN(x * RN(1/c -
 * RN(1/c)))) the correctly rounded x / c for EVERY normal x?  Decided per divisor by trying the finitely many operands whose
 * quotient lies within the sequence's error of a rounding midpoint (fast_common.cuh: div2_exact) -- a pure host function.
 * The sweep uses that sequence only when all four of its divisors (G dt, eta, eta_ve, 3.0) pass, the four-operation sequence
 * otherwise; chmy_last_division_mode: what the last fused 3D sweep of the context used (0: four operations, 1: div.rn.f64,
 * 2: two operations).  Env CHMY_DIV2=0 keeps the four-operation sequence. */
int chmy_division_two_op_exact(double c, int32_t* exact);
int chmy_selftest_division2(chmy_ctx* ctx, double c, long long n, unsigned long long seed, unsigned long long* mismatches, int* proved);
int chmy_last_division_mode(const chmy_ctx* ctx, int32_t* mode);
/* -1 keeps a setting.  disable_fast_kernels: run every op with the generic one-thread-per-cell kernels (A/B
 * parity of the tuned kernels); force_true_division: div.rn.f64 everywhere.  Env: CHMY_NO_FAST=1, CHMY_TRUE_DIV=1. */
int chmy_set_tuning(int disable_fast_kernels, int force_true_division);
/* Launches with boundary batches (KernelLaunch.jl:152-183).  overlap (-1 keeps): 1 (default) = the batches and the halo
 * exchange overlap the kernel -- the reference's inner region + slabs on two streams for the plain kernels (only when a
 * side is Connected: without a neighbour six extra launches buy nothing), boundary tiles first + a retire counter for the
 * fused 3D sweep (one launch; the boundary stream sleeps on the counter, then runs the batches while the interior tiles are
 * still computing -- again only when a side is Connected; 2 = also without a neighbour, for tests and A/B); 0 = one kernel,
 * then the batches, on one stream.  outer_width is a hint; results are identical and
 * the order is deterministic (no run-time tuning: every rank of a topology takes the same one).
 * bc_fold (-1 keeps): 1 (default) = a batch set without exchange runs as ONE launch for all dimensions, sides and fields
 * (same bits as the reference's D = N..1 order, batch.jl:20-29); 0 = one launch per dimension.
 * Env: CHMY_OVERLAP=0|1, CHMY_BC_FOLD=0|1. */
int chmy_set_launch_tuning(chmy_ctx* ctx, int overlap, int bc_fold);
int chmy_overlapped_count(const chmy_ctx* ctx, uint64_t* launches);   /* launches whose batches ran behind a still-running sweep */
/* The split decision of a launch with boundary batches, without launching: *split = 0 -> one full-range kernel followed by
 * the batches; 1 -> inner region [wl, n+2-wr) per dim on the main stream and, for D = N..1, the two slabs of widths
 * wl[D] / wr[D] on the boundary stream (KernelLaunch.jl:63-87 with outer_width replaced by wl / wr).  pref: the slab widths
 * the op's kernel prefers ({60, 0, 0} for the 2D sweeps) or NULL; overlap: the context's policy above.  Pure function. */
int chmy_launch_split_plan(const chmy_launch_desc* desc, const int32_t* pref, int32_t overlap, int32_t* split, int32_t wl[3], int32_t wr[3]);
/* the launch order of the fused sweep's tiles: out[4*c .. 4*c+3] = (bx, by, bz, is_boundary) of linear cluster index c, for
 * g[] tiles per dim with interior index ranges [i0, i1); tail = 1: the order of an overlapped launch (every boundary tile
 * retires before the last stretch of interior tiles), 0: natural order -- a pure function, for the CPU tests */
int chmy_selftest_tile_order(const int32_t g[3], const int32_t i0[3], const int32_t i1[3], int32_t tail, int32_t* out);

/* ---- lazily fused update_stress! -> update_velocity! (cross-launch fusion) -----------------
 * The reference runs the two kernels of a PT iteration as two `launch` calls (stokes_3d_inc_ve_T.jl:163-165); the
 * velocity kernel re-reads what the stress kernel has just written.  With fusion enabled a 3D
 * `launch(update_stress!)` without bc is deferred, and if the next call on the context is the matching
 * `launch(update_velocity!; bc)` both run as one sweep that keeps the new stresses on chip (30 instead of 40 array
 * passes).  Any other call executes the deferred launch first, so results are identical with and without fusion.
 * The sweep writes tau, Pr and V into shadow buffers that are swapped with the fields' storage afterwards:
 * pointers obtained from chmy_field_get_info are invalidated by a fused launch (PITCHED layout only; DENSE fields
 * and anything the sweep cannot handle fall back to the two kernels).
 * `enable`: 0 = off; 1 = the 3D pair above; 3 = additionally the flux -> update pairs below (same deferred-launch
 * protocol, same results; all measured and bit-exact on B200, profiles/r2_c1_*):
 *   update_stress! + update_velocity! 2D            stokes_2d_inc_ve_T.jl:146-147   24 -> 18 array passes
 *   compute_q! + update_C!                          diffusion_2d_perf.jl:28-29       7 ->  4
 *   update_thermal_flux! + update_thermal! 2D       stokes_2d_inc_ve_T.jl:151-152    9 ->  7
 *   update_thermal_flux! + update_thermal! 3D       stokes_3d_inc_ve_T.jl:167-168   12 ->  9                    */
int chmy_set_fusion(chmy_ctx* ctx, int enable);
int chmy_fused_count(const chmy_ctx* ctx, uint64_t* sweeps);          /* fused sweeps launched so far            */
/* pairs that ran as two kernels because the device had no memory for their shadow buffers (the sweeps write through
 * ping-pong twins of the fields they update): results are the same, the pair just moves 40 instead of 30 passes */
int chmy_fusion_fallback_count(const chmy_ctx* ctx, uint64_t* pairs);
/* Per-context tile geometry of the fused 3D sweep: rows of a CTA (4|6|8), CTAs per thread-block cluster along y (1..8), planes
 * per z-chunk (0 keeps a setting); variant: bit 0 = relaxed cluster-barrier arrive behind a CTA-scope fence instead of the
 * release arrive (-1 keeps).  Defaults 6, 4, 64, 1 = the measured optimum at 767^3 (profiles/README.md).
 * Env (read when the context is created): CHMY_FUSE_TYB, CHMY_FUSE_CL, CHMY_FUSE_CZ, CHMY_FUSE_VARIANT. */
int chmy_set_fused_tuning(chmy_ctx* ctx, int rows_per_cta, int cluster_size, int z_chunk, int variant);
/* 2D sweeps: rows per y-chunk of a warp; rows whose operands are requested ahead of the arithmetic (1|2|4, flux pairs);
 * 3D thermal sweep: planes per z-chunk (0 keeps a setting).  Env: CHMY_FUSE2D_CY, CHMY_FUSE2D_UNROLL, CHMY_FUSE_T3_CZ. */
int chmy_set_fused2d_tuning(chmy_ctx* ctx, int rows_per_chunk, int unroll, int thermal3_planes_per_chunk);

/* halo slab pack/unpack exposed for bit-exact parity tests of src/Distributed/communication_views.jl:1-34 */
int chmy_halo_slab_len(const chmy_field* f, int dim, int64_t* len);
int chmy_halo_pack(chmy_ctx* ctx, const chmy_field* f, int dim, int side, void* host_buf);      /* elements of the field's type */
int chmy_halo_unpack(chmy_ctx* ctx, chmy_field* f, int dim, int side, const void* host_buf);

#ifdef __cplusplus
}
#endif
#endif /* CHMY_B200_H */
