/*
 * chmy_oracle.h -- CPU restatement of the Chmy.jl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle: a plain-C restatement of the arithmetic that the reference
 * (PTsolvers/Chmy.jl v0.1.25, /root/reference) performs on its KernelAbstractions CPU
 * backend for the staggered-grid PT diffusion / Stokes(+T) kernels, the Dirichlet/Neumann
 * boundary batches and the halo pack/unpack views.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the shipped product
 * (chmy.jl_b200/, libchmy_b200.so) never links, imports or calls anything in this directory.
 *
 * Pinning: the reference is pure Julia and no Julia toolchain exists in this image, so the
 * reference itself cannot be executed.  The oracle is pinned against every known-answer the
 * reference's own tests hold for this path (test/test_grids.jl, test_fields.jl,
 * test_boundary_conditions.jl, test_grid_operators.jl, test_interpolations.jl) -- see
 * tests/test_oracle_golden.py.  Halo exchange, Launcher splitting and the example solvers have
 * NO golden vectors in the reference ("parity unpinned" for those rows); they
 * are checked through invariants (split == unsplit, N ranks == 1 rank, pack/unpack round trip); the solver kernels'
 * flattened arithmetic is additionally pinned bit-for-bit on a literal transliteration of the reference's kernel source
 * over the pinned point operators (tests/test_oracle_transliteration.py).
 *
 * Conventions (src/Fields/field.jl:6-22,56-62): a Field of logical size d[] is stored as a dense
 * column-major array of d[]+4 elements per active dimension; logical index I (1-based, as in the
 * reference) lives at storage offset I+1 (0-based), i.e. I=-1 and I=d+2 are zero padding,
 * I=0 and I=d+1 the halo.  Every +,-,*,/ is one IEEE binary64 operation in the written order;
 * fma() appears only where the reference writes muladd.  Build with -ffp-contract=off.
 */
#ifndef CHMY_ORACLE_H
#define CHMY_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Element type.  The reference instantiates everything for Float32 and Float64 (test/common.jl:9).  This file is built
 * twice: libchmy_oracle.so (og_real = double, the solvers' type) and libchmy_oracle_f32.so (-DOG_F32: og_real = float)
 * for the Float32 rows of the reference's tests (grids, fields, boundary conditions, operators, halo views).  In the
 * Float32 build every operation is a binary32 operation where the reference's would be; the example-solver ops keep
 * Float64 literals in the reference (2.0, 3.0, 0.5: they would promote), so they are Float64-only here as well. */
/* Element type of grids and fields.  The example kernels contain Float64 literals (0.5, 2.0, 3.0, 0.0): in a Float32 run
 * Julia promotes exactly the sub-expressions that meet one of them to Float64 and rounds once when the result is stored
 * into the Float32 array (stokes_3d_inc_ve_T.jl:27-45,62-70).  og_wide is the type of those sub-expressions; C's usual
 * arithmetic conversions then reproduce Julia's promotion operation by operation (FLT_EVAL_METHOD == 0). */
typedef double og_wide;
#ifdef OG_F32
typedef float og_real;
#else
typedef double og_real;
#endif

enum { OG_CENTER = 0, OG_VERTEX = 1 };
enum { OG_BOUNDED = 0, OG_CONNECTED = 1 };
enum { OG_DIRICHLET = 0, OG_NEUMANN = 1 };

/* UniformGrid: src/Grids/uniform_axis.jl:1-12, structured_grid.jl:27-39 */
typedef struct {
    int32_t nd;
    int64_t n[3];              /* number of cells (centers) per dim; 1 for inactive dims */
    og_real  origin[3];
    og_real  extent[3];
    og_real  spacing[3];        /* extent / n            (uniform_axis.jl:8) */
    og_real  inv_spacing[3];    /* inv(spacing)          (uniform_axis.jl:9) */
    int32_t conn[3][2];
} og_grid;

/* Field{T,N,L,H=1}: src/Fields/field.jl:6-22 */
typedef struct {
    int32_t nd;
    int32_t loc[3];
    int64_t d[3];              /* logical dims (size(grid, loc)); 1 for inactive dims */
    int64_t sd[3];             /* storage dims d+4 (1 for inactive dims) */
    int64_t o[3];              /* storage offset of logical index 0: 1 for active dims (I+1), else 0 */
    og_real* data;
} og_field;

/* FunctionField with the `init_incl` body used by the Stokes drivers
 * (examples/stokes_3d_inc_ve_T_mpi_perf.jl:141-142, src/Fields/function_field.jl:49-59) */
typedef struct {
    int32_t active;            /* 0: rho_g comes from a stored field */
    int32_t loc[3];
    og_real  c0[3];             /* x0,y0,z0 */
    og_real  r, in, out;
} og_inclusion;

void   og_grid_init(og_grid* g, int nd, const int64_t* n, const og_real* origin, const og_real* extent);
og_real og_coord(const og_grid* g, int dim, int loc, int64_t i);
void   og_field_init(og_field* f, const og_grid* g, const int32_t* loc, og_real* data);
int64_t og_field_storage_len(const og_grid* g, const int32_t* loc);

void   og_set_inclusion(const og_grid* g, og_field* f, const og_inclusion* inc);
og_real og_maxabs_interior(const og_field* f);

/* region boxes are inclusive logical index ranges lo[d]..hi[d] */
void og_compute_q(const og_grid* g, og_field* qx, og_field* qy, const og_field* C, og_real chi,
                  const int64_t* lo, const int64_t* hi);
void og_update_C(const og_grid* g, og_field* C, const og_field* qx, const og_field* qy, og_real dt,
                 const int64_t* lo, const int64_t* hi);
void og_update_old(const og_grid* g, int npairs, og_field* const* dst, const og_field* const* src,
                   const int64_t* lo, const int64_t* hi);
void og_update_stress2(const og_grid* g, og_field* const* tau, og_field* Pr, og_field* divV,
                       const og_field* const* V, const og_field* const* tau_old,
                       og_real eta, og_real eta_ve, og_real G, og_real dt, og_real dtau_Pr, og_real dtau_r,
                       const int64_t* lo, const int64_t* hi);
void og_update_stress3(const og_grid* g, og_field* const* tau, og_field* Pr, og_field* divV,
                       const og_field* const* V, const og_field* const* tau_old,
                       og_real eta, og_real eta_ve, og_real G, og_real dt, og_real dtau_Pr, og_real dtau_r,
                       const int64_t* lo, const int64_t* hi);
void og_update_velocity2(const og_grid* g, og_field* const* V, og_field* const* rV, const og_field* Pr,
                         const og_field* const* tau, const og_field* rhog, const og_inclusion* inc,
                         og_real eta_ve, og_real nudtau, const int64_t* lo, const int64_t* hi);
void og_update_velocity3(const og_grid* g, og_field* const* V, og_field* const* rV, const og_field* Pr,
                         const og_field* const* tau, const og_field* rhog, const og_inclusion* inc,
                         og_real eta_ve, og_real nudtau, const int64_t* lo, const int64_t* hi);
void og_update_thermal_flux(const og_grid* g, og_field* const* qT, const og_field* T, const og_field* const* V,
                            og_real lambda, const int64_t* lo, const int64_t* hi);
void og_update_thermal(const og_grid* g, og_field* T, const og_field* T_old, const og_field* const* qT,
                       og_real dt, const int64_t* lo, const int64_t* hi);

/* one (dim, side) of one field: src/BoundaryConditions/first_order_boundary_condition.jl:34-84,
 * face range from batch.jl:159-184 */
void og_bc_apply(const og_grid* g, og_field* f, int dim, int side, int kind, og_real value);
/* Field-valued condition: vf is a lower-dimensional field read at remove_dim(dim, I) (:38-40) */
void og_bc_apply_field(const og_grid* g, og_field* f, int dim, int side, int kind, og_real value, const og_field* vf);

/* halo slabs: src/Distributed/communication_views.jl:1-34 */
int64_t og_slab_len(const og_field* f, int dim);
void og_pack_send(const og_field* f, int dim, int side, og_real* buf);
void og_unpack_recv(og_field* f, int dim, int side, const og_real* buf);

/* generic operators used by the reference's operator tests (test/test_grid_operators.jl,
 * test/test_interpolations.jl): src/GridOperators/partial_derivatives.jl, interpolation.jl */
og_real og_partial(const og_grid* g, const og_field* f, int dim, int64_t i, int64_t j, int64_t k);
og_real og_partial2(const og_grid* g, const og_field* f, int dim, int64_t i, int64_t j, int64_t k);
og_real og_lerp(const og_grid* g, const og_field* f, const int32_t* to, int64_t i, int64_t j, int64_t k);
og_real og_dkd(const og_grid* g, const og_field* f, const og_field* kf, int dim, int64_t i, int64_t j, int64_t k);
og_real og_hlerp(const og_grid* g, const og_field* f, const int32_t* to, int64_t i, int64_t j, int64_t k);
/* dst[I] = OP(src...)[I] over the box [lo, hi]; kind as documented at the definition (field_operators.jl:2-121) */
void og_apply_operator(const og_grid* g, int kind, int dim, og_field* const* dst, const og_field* const* src,
                       const og_field* kf, const int64_t* lo, const int64_t* hi);

int og_num_threads(void);
void og_set_num_threads(int n);
int og_real_bytes(void);   /* sizeof(og_real): 8 | 4 */

#ifdef __cplusplus
}
#endif
#endif
