// fused_pairs2d.cuh -- EXPERIMENTAL (round-2 candidates; opt-in through chmy_set_fusion(ctx, 3); proven by the host
// emulation, not yet run on a GPU): the two flux -> update pairs of the 2D drivers as ONE y-marching sweep each,
//   KIND 0  compute_q! + update_C!                  (examples/diffusion_2d.jl:8-19, diffusion_2d_perf.jl:25-32)
//   KIND 1  update_thermal_flux! + update_thermal!  (examples/stokes_2d_inc_ve_T.jl:45-60)
// Both have the shape of the Stokes pair (fused_sv2d.cuh): the first launch writes fluxes on the op's whole index range
// [0, n+1]^2, the second one differences them.  Two kernels move R1+W2 and R3+W1 = 7 array passes (diffusion) or
// R3+W2 and R3+W1 = 9 (thermal); one sweep that keeps the new fluxes in registers moves R1+W3 = 4 or R4+W3 = 7.
//
// A lane owns 2 x-adjacent cells, a warp covers 64 cells of which lanes 1..30 are interior (lane 0 only supplies
// C[i-1] to lane 1, lane 31 only recomputes the flux q.x[i+2] its left neighbour differences), and the warp marches
// along y: iteration jp computes the fluxes of row jp (phase A) and then updates row jp-1 (phase B) from the carried
// fluxes of row jp-1, phase A's row jp and the right lane's first flux (one warp shuffle).  No shared memory, no
// barrier.  The updated field (C or T) is read from its current buffer and written to its shadow buffer (ping-pong:
// a neighbouring tile still needs the old values for its own halo fluxes); the fluxes are written in place (nobody
// reads old fluxes).  Outside the op's index range the first kernel never ran, so there the "new" flux is the stored
// one.  Arithmetic order is that of ops.cu / the reference.
//
// Compiles with nvcc (ops_fused2d.cu) and with a host compiler (tests/emul/fused_emul2d.cpp runs the 32 lanes of a
// warp in lock-step; the shuffle of phase B is the only line that differs).
#pragma once
#include "fused_sv.cuh"   // d2, ld2/st2, fsv_from_left, FSV_LANES, FSV_XI, FHD

#ifndef __CUDACC__
static inline double jl_max0(double v) { return (v != v) ? v : std::fmax(v, 0.0); }   // Julia Base.max(v, 0.0)
static inline double jl_min0(double v) { return (v != v) ? v : std::fmin(v, 0.0); }
#endif

struct FusedQ2P {
    const double* Cc;        // current C (KIND 0) / T (KIND 1)
    double*       Cn;        // its shadow buffer
    const double* base;      // KIND 1: T_old ; KIND 0: unused
    double *      qx, *qy;   // fluxes, written in place
    const double *Vx, *Vy;   // KIND 1: advecting velocity
    int s_cc, s_vc, s_cv;    // row strides: CC C T T_old ; VC q.x V.x ; CV q.y V.y
    int lo[2], hi[2];        // update / store box, hi exclusive
    int flo[2], fhi[2];      // index range of the op
    double idx, idy, coef, dt;   // coef: chi (KIND 0) / lambda (KIND 1)
    int cy;                  // rows per y-chunk
};

struct FusedQ2T {
    int  lane, i, j0, j1;
    bool s_act;              // this lane loads and computes fluxes
    int  nv;                 // cells of the pair that are updated and stored (0, 1 or 2)
    bool fx0, fx1;
    long long cc, vc, cv;    // element offsets of (i, jp)
    d2 c_jm, c_j;            // C(jp-1) ; phase A -> B: C(jp)
    d2 qxC, qyC;             // new fluxes of row jp-1
};

// seg = row-segment index along x, cyc = y-chunk index; the march starts at row j0 (row j0-1 only supplies C)
FHD void fq2_init(FusedQ2T& s, const FusedQ2P& p, int lane, int seg, int cyc) {
    s.lane = lane;
    s.i  = p.lo[0] - 2 + seg * FSV_XI + 2 * lane;
    s.j0 = p.lo[1] + cyc * p.cy;
    s.j1 = s.j0 + p.cy < p.hi[1] ? s.j0 + p.cy : p.hi[1];
    s.s_act = s.i <= p.hi[0];
    int nv = p.hi[0] - s.i;
    nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
    s.nv = (lane >= 1 && lane <= FSV_LANES - 2) ? nv : 0;
    s.fx0 = s.i >= p.flo[0] && s.i < p.fhi[0];
    s.fx1 = s.i + 1 >= p.flo[0] && s.i + 1 < p.fhi[0];
    s.cc = (long long)s.i + (long long)s.j0 * p.s_cc;
    s.vc = (long long)s.i + (long long)s.j0 * p.s_vc;
    s.cv = (long long)s.i + (long long)s.j0 * p.s_cv;
    const d2 z = fsv_zero();
    s.c_jm = s.c_j = s.qxC = s.qyC = z;
    if (s.s_act) s.c_jm = ld2(p.Cc + s.cc - p.s_cc);
}

// operands of the row `ahead` rows above the thread's current row (s.cc ... address row jp): the kernel requests the rows
// of a whole group of iterations before it computes the first one, so that several loads per thread are in flight
struct FusedQ2L {
    d2 c, vx, vy;
};
template <int KIND>
FHD void fq2_load(const FusedQ2T& s, const FusedQ2P& p, int ahead, FusedQ2L& L) {
    const d2 z2 = fsv_zero();
    L.c = z2; L.vx = z2; L.vy = z2;
    if (s.s_act) {
        L.c = ld2(p.Cc + s.cc + (long long)ahead * p.s_cc);
        if (KIND == 1) {
            L.vx = ld2(p.Vx + s.vc + (long long)ahead * p.s_vc);
            L.vy = ld2(p.Vy + s.cv + (long long)ahead * p.s_cv);
        }
    }
}

// ---- phase A: fluxes of row jp from its preloaded operands -> sn[0] = q.x pair, sn[1] = q.y pair; stores for the cells
// this thread owns
template <int KIND>
FHD void fq2_phase_a(FusedQ2T& s, const FusedQ2P& p, int jp, const FusedQ2L& L, d2 sn[2]) {
    const d2 z2 = fsv_zero();
    const d2 c = L.c, vx = L.vx, vy = L.vy;
    const bool okl = s.s_act && s.lane > 0;
    const double c_im1 = fsv_from_left(c.y, p.Cc + s.cc - 1, okl);   // lane 0 gets a don't-care value (its fluxes are unused)
    const bool fy = jp >= p.flo[1] && jp < p.fhi[1];
    d2 qx = z2, qy = z2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double a_c = h ? c.y : c.x, a_cim = h ? c.x : c_im1, a_cjm = h ? s.c_jm.y : s.c_jm.x;
        double fx, fy_;
        if (KIND == 0) {   // diffusion_2d.jl:10-11
            fx  = (-p.coef) * ((a_c - a_cim) * p.idx);
            fy_ = (-p.coef) * ((a_c - a_cjm) * p.idy);
        } else {           // stokes_2d_inc_ve_T.jl:47-53
            const double a_vx = h ? vx.y : vx.x, a_vy = h ? vy.y : vy.x;
            fx  = ((-p.coef) * ((a_c - a_cim) * p.idx) + jl_max0(a_vx) * a_cim) + jl_min0(a_vx) * a_c;
            fy_ = ((-p.coef) * ((a_c - a_cjm) * p.idy) + jl_max0(a_vy) * a_cjm) + jl_min0(a_vy) * a_c;
        }
        // outside the op's index range the flux kernel never ran: the update sees the stored flux
        const bool in = (h ? s.fx1 : s.fx0) && fy;
        if (!in) {
            fx  = s.s_act ? p.qx[s.vc + h] : 0.0;
            fy_ = s.s_act ? p.qy[s.cv + h] : 0.0;
        }
        if (h) { qx.y = fx; qy.y = fy_; } else { qx.x = fx; qy.x = fy_; }
    }
    if (jp >= s.j0 && jp < s.j1) {
        if (s.nv == 2) {
            st2(p.qx + s.vc, qx); st2(p.qy + s.cv, qy);
        } else if (s.nv == 1) {
            p.qx[s.vc] = qx.x; p.qy[s.cv] = qy.x;
        }
    }
    sn[0] = qx; sn[1] = qy;
    s.c_j = c;
}

// ---- phase B: update row j = jp-1 from the carried fluxes of row jp-1, phase A's q.y of row jp and qx_ip2 = the right
// lane's first q.x of row jp-1 (warp shuffle on the device); then rotate the carried rows.
template <int KIND>
FHD void fq2_phase_b(FusedQ2T& s, const FusedQ2P& p, int jp, const d2 sn[2], double qx_ip2) {
    if (s.nv > 0 && jp >= s.j0 + 1) {
        const long long cc = s.cc - p.s_cc;
        d2 b;
        if (KIND == 0) b = s.c_jm;
        else if (s.nv == 2) b = ld2(p.base + cc);
        else { b.x = p.base[cc]; b.y = 0.0; }
        d2 r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a_qx = h ? s.qxC.y : s.qxC.x, a_qxip = h ? qx_ip2 : s.qxC.y;
            const double a_qy = h ? s.qyC.y : s.qyC.x, a_qyjp = h ? sn[1].y : sn[1].x;
            const double dv = (a_qxip - a_qx) * p.idx + (a_qyjp - a_qy) * p.idy;   // divg, field_operators.jl:50-55
            const double v  = (h ? b.y : b.x) - p.dt * dv;
            if (h) r.y = v; else r.x = v;
        }
        if (s.nv == 2) st2(p.Cn + cc, r); else p.Cn[cc] = r.x;
    }
    s.qxC = sn[0]; s.qyC = sn[1];
    s.c_jm = s.c_j;
    s.cc += p.s_cc; s.vc += p.s_vc; s.cv += p.s_cv;
}
