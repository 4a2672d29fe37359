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
