// bc_kernels.cuh -- per-point bodies of the boundary-condition and halo-slab kernels (bc.cu), as plain C++ shared by nvcc
// and the host compiler: tests/emul/bc_emul.cpp runs them point by point on the CPU and tests/test_bc_emulation.py proves
// them bit-identical to the oracle in both element types (Float64 | Float32, test/common.jl:9) without a GPU.
//
// Boundary rules: src/BoundaryConditions/first_order_boundary_condition.jl:34-84 on a uniform grid
//   Dirichlet, field Vertex along D : f[b] = v                          b  = 1 | d      (the boundary node)
//   Dirichlet, field Center along D : f[h] = muladd(2, v - f[nb], f[nb])  h = 0 | d+1 ; nb = 1 | d
//   Neumann  , any location         : f[h] = muladd(spacing_D, -/+q, f[nb])
// Face range: src/BoundaryConditions/batch.jl:159-184 -- transverse index I_t = J-1 in 0..n_t+2, fields of a batch
// applied in batch order.  Halo slabs: src/Distributed/communication_views.jl:1-34.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define BCK_HD __host__ __device__ __forceinline__
#else
#define BCK_HD inline
#endif

#define BCK_MAX_FIELDS 8      // == CHMY_MAX_BATCH_FIELDS
#define BCK_DIRICHLET 0       // == CHMY_DIRICHLET

// View of a Field: p addresses logical index 0 of every active dimension, x stride 1 (same layout as FVT<T>, common.cuh)
template <class T>
struct BckView {
    T*        p;
    long long sy, sz;
};
template <class T>
BCK_HD T bck_ld(const BckView<T>& f, int i, int j, int k) { return f.p[(long long)i + (long long)j * f.sy + (long long)k * f.sz]; }
template <class T>
BCK_HD void bck_st(const BckView<T>& f, int i, int j, int k, T v) { f.p[(long long)i + (long long)j * f.sy + (long long)k * f.sz] = v; }

template <class T>
struct BcEntry {
    BckView<T> f;
    int    kind;      // chmy_bc_kind
    int    vertex;    // location of the field along the BC dim
    int    d;         // logical size of the field along the BC dim
    int    side;      // 0 | 1
    T      value;
    const T*  vp;     // Field-valued condition: logical (0[,0]) of the (N-1)-dimensional value field, else nullptr
    long long vsy;
};

template <class T>
struct BcBatchDev {
    int        n;                                // entries (both sides of one dim)
    int        dim;
    int        nt[2];                            // transverse extents (n_t + 3 points each; 1 when absent)
    T          spacing;
    BcEntry<T> e[2 * BCK_MAX_FIELDS];
};

// One face point (a, c) = the two transverse indices; both sides and all fields of a dimension.  Entries of different
// sides touch disjoint cells and different fields are independent, so the reference's sequential order (side 1 then 2,
// fields in batch order) is preserved per cell.
template <class T>
BCK_HD void bc_point(const BcBatchDev<T>& b, int a, int c) {
    for (int q = 0; q < b.n; ++q) {
        const BcEntry<T>& e = b.e[q];
        int I[3], N[3];
        // insert_dim(dim, (a, c), idx)  -- src/utils.jl:47-51
        int t = 0;
        const int tr[2] = {a, c};
        const int bnode = e.side == 0 ? 1 : e.d;
        const int hnode = e.side == 0 ? 0 : e.d + 1;
        for (int dd = 0; dd < 3; ++dd) {
            if (dd == b.dim) { I[dd] = hnode; N[dd] = bnode; }
            else { I[dd] = N[dd] = (t < 2 ? tr[t] : 0); ++t; }
        }
        // value(bc, grid, loc, dim, I...): Number | bc.value[remove_dim(dim, I)...]  (first_order_boundary_condition.jl:34-40)
        const T val = e.vp ? e.vp[(long long)a + (long long)c * e.vsy] : e.value;
        if (e.kind == BCK_DIRICHLET) {
            if (e.vertex) {
                bck_st(e.f, N[0], N[1], N[2], val);
            } else {
                const T nb = bck_ld(e.f, N[0], N[1], N[2]);
                bck_st(e.f, I[0], I[1], I[2], (T)fma((T)2.0, val - nb, nb));
            }
        } else {
            const T qs = e.side == 0 ? -val : val;
            bck_st(e.f, I[0], I[1], I[2], (T)fma(b.spacing, qs, bck_ld(e.f, N[0], N[1], N[2])));
        }
    }
}

// ---------------------------------------------------------------------------------------------- all dimensions at once
// bc!(arch, grid, batchset) applies the batches dimension by dimension, D = N..1 (batch.jl:20-29), and a later dimension
// reads cells an earlier one wrote (faces span the transverse range 0..n_t+2, edges and corners included) -- which is why
// the per-dimension kernel needs one launch per dimension.  Every rule, though, is a pure function of ONE cell:
//     f[target] = rule(value, f[neighbour])          (or just `value` for Dirichlet on a Vertex location)
// so the final content of a cell is a chain of at most N rules ending in a cell NO batch writes.  bc_all_point evaluates
// that chain for one target cell straight from the untouched cells -- every target is computed independently, nothing
// written by the kernel is read by it, and all dimensions, sides and fields run as ONE launch with the same bits as
// the sequential order (tests/test_bc_emulation.py: bit-identical to the oracle's dimension-by-dimension bc!).
template <class T>
struct BcRule {
    int       kind;       // -1: no condition on this (dim, side) ; else chmy_bc_kind
    T         value;
    const T*  vp;         // Field-valued condition: logical (0[,0]) of the (N-1)-dimensional value field, else nullptr
    long long vsy;
};
template <class T>
struct BcAllField {
    BckView<T> f;
    int        d[3];          // logical size per dim (1 for inactive dims)
    int        vertex[3];     // location per dim
    BcRule<T>  r[3][2];
};
#define BCK_ALL_FIELDS 6
template <class T>
struct BcAllDev {
    int           nf, nd;
    int           n[3];       // grid cells per dim: face points span 0..n+2 in every transverse dim
    T             spacing[3];
    BcAllField<T> fld[BCK_ALL_FIELDS];
    int           nact;                                   // faces that carry a condition ...
    unsigned char act[6 * BCK_ALL_FIELDS];                // ... as (field * 3 + dim) * 2 + side: one slice of the launch grid each
};

template <class T>
BCK_HD int bc_all_target(const BcAllField<T>& F, int D, int s) {
    const bool node = F.r[D][s].kind == BCK_DIRICHLET && F.vertex[D];
    return s == 0 ? (node ? 1 : 0) : (node ? F.d[D] : F.d[D] + 1);
}
// the side of dim D whose condition writes cell I (or -1): I[D] is that side's target and every other index is a face point
template <class T>
BCK_HD int bc_all_side(const BcAllDev<T>& b, const BcAllField<T>& F, int D, const int I[3]) {
    for (int t = 0; t < b.nd; ++t)
        if (t != D && (I[t] < 0 || I[t] > b.n[t] + 2)) return -1;
    for (int s = 0; s < 2; ++s)
        if (F.r[D][s].kind >= 0 && I[D] == bc_all_target(F, D, s)) return s;
    return -1;
}
template <class T>
BCK_HD T bc_all_value(const BcAllDev<T>& b, const BcRule<T>& r, int D, const int I[3]) {
    if (!r.vp) return r.value;
    int tr[2] = {0, 0}, t = 0;
    for (int a = 0; a < b.nd; ++a)
        if (a != D) tr[t++] = I[a];                    // remove_dim(dim, I)
    return r.vp[(long long)tr[0] + (long long)tr[1] * r.vsy];
}

// face point (a, c) of (field q, dim D, side s): computes the FINAL content of its target cell unless a dimension that is
// applied later (D' < D) writes that cell too (then that dimension's thread computes it)
template <class T>
BCK_HD void bc_all_point(const BcAllDev<T>& b, int q, int D, int s, int a, int c) {
    const BcAllField<T>& F = b.fld[q];
    if (F.r[D][s].kind < 0) return;
    int I[3] = {0, 0, 0};
    {
        const int tr[2] = {a, c};
        int t = 0;
        for (int dd = 0; dd < b.nd; ++dd) I[dd] = dd == D ? bc_all_target(F, D, s) : tr[t++];
    }
    if (bc_all_side(b, F, D, I) != s) return;          // the other side's target coincides (degenerate sizes): it wins below
    for (int Dl = 0; Dl < D; ++Dl)
        if (bc_all_side(b, F, Dl, I) >= 0) return;
    // walk back through the dimensions in reverse order of application (D, D+1, ...: later applied first)
    int   cur[3] = {I[0], I[1], I[2]};
    int   cd[3], cs[3], nc = 0;
    T     cv[3];
    bool  terminal = false;
    T     v = (T)0;
    for (int De = D; De < b.nd; ++De) {
        const int se = De == D ? s : bc_all_side(b, F, De, cur);
        if (se < 0) continue;
        const BcRule<T>& r = F.r[De][se];
        const T val = bc_all_value(b, r, De, cur);
        if (r.kind == BCK_DIRICHLET && F.vertex[De]) { v = val; terminal = true; break; }
        cd[nc] = De; cs[nc] = se; cv[nc] = val; ++nc;
        cur[De] = se == 0 ? 1 : F.d[De];               // the neighbour the rule reads
    }
    if (!terminal) v = bck_ld(F.f, cur[0], cur[1], cur[2]);
    for (int k = nc - 1; k >= 0; --k) {
        const BcRule<T>& r = F.r[cd[k]][cs[k]];
        if (r.kind == BCK_DIRICHLET) v = (T)fma((T)2.0, cv[k] - v, v);
        else v = (T)fma(b.spacing[cd[k]], cs[k] == 0 ? -cv[k] : cv[k], v);
    }
    bck_st(F.f, I[0], I[1], I[2], v);
}

// ---------------------------------------------------------------------------------------------- halo slabs
// send index: side 1 -> 1+overlap, side 2 -> d-overlap (overlap = 1 for Vertex, 0 for Center);
// recv index: side 1 -> 0, side 2 -> d+1; every other dimension spans the whole padded extent -1..d+2.
template <class T>
struct SlabEntry {
    BckView<T> f;
    int       idx;        // logical index of the slab along dim
    int       e0, e1;     // transverse extents (sd_t), 1 when absent
    long long off;        // element offset of this field's slab in the buffer
};
template <class T>
struct SlabBatch {
    int          n, dim, nd;
    SlabEntry<T> e[BCK_MAX_FIELDS];
};

// element (a, c) of the slab of field `q`: storage transverse indices, column-major in the message
template <bool PACK, class T>
BCK_HD void slab_point(const SlabBatch<T>& b, T* __restrict__ buf, int q, int a, int c) {
    const SlabEntry<T>& e = b.e[q];
    if (a >= e.e0 || c >= e.e1) return;
    int I[3], t = 0;
    const int tr[2] = {a - 1, c - 1};           // storage 0 <-> logical -1
    for (int dd = 0; dd < 3; ++dd) {
        if (dd == b.dim) I[dd] = e.idx;
        else if (dd >= b.nd) I[dd] = 0;          // inactive dimension
        else { I[dd] = tr[t]; ++t; }
    }
    const long long p = e.off + (long long)a + (long long)c * e.e0;
    if (PACK) buf[p] = bck_ld(e.f, I[0], I[1], I[2]);
    else bck_st(e.f, I[0], I[1], I[2], buf[p]);
}
