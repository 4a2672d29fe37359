// fused_sv.cuh -- update_stress! + update_velocity! of the 3D Stokes PT iteration in ONE z-marching sweep
// (examples/stokes_3d_inc_ve_T.jl:23-57; SURVEY.md §8(f) row 4 "cross-launch fusion").
//
// Why: the two-kernel formulation moves 40 array passes per iteration (stress R16+W8, velocity R10+W6); the
// velocity kernel re-reads Pr and the six stress components the stress kernel has just written, and V is read twice.
// One sweep that keeps the freshly computed stresses on chip moves R16 + W14 = 30 passes.
//
// How (per CTA = 32 lanes x TYB warp-rows, optionally CL CTAs stacked along y in a thread-block cluster):
//   * a lane owns 2 x-adjacent cells (128-bit accesses, as in ops_fast.cu); a warp row covers 64 cells of which the
//     middle 60 (lanes 1..30) are "interior": lanes 0 and 31 only recompute the stresses the interior needs from
//     the neighbouring row segment, so no x-neighbour of a NEW value ever crosses a warp;
//   * the first and last row of the cluster are halo rows in the same sense along y; CTAs of a cluster read the
//     boundary rows of their neighbours through distributed shared memory;
//   * iteration kp of the z-march computes the stresses of plane kp (phase A: global loads, arithmetic, stores of
//     the new tau / Pr / divV), publishes them in shared memory (double-buffered by plane parity), and then
//     updates the velocity of plane kp-1 (phase B), whose x/y stress neighbours come from shared memory, whose
//     k+1 neighbours are the values of phase A still in registers and whose k-1 neighbours are carried registers.
//     One (split arrive/wait) barrier per plane;
//   * in-place updates would race with the halo recomputation of neighbouring tiles, so tau, Pr and V are read from
//     the fields' current buffers and written to their shadow buffers (ping-pong, swapped by the host after the
//     launch).  Halo cells outside the op's index range [0, n+1]^3 are never produced by update_stress! in the
//     reference, so there the "new" value is the stored one.
//
// Arithmetic order is exactly that of ops.cu / ops_fast.cu / the reference; the file compiles both with nvcc (the
// kernel in ops_fused.cu) and with a host compiler (tests/emul/fused_emul.cpp executes the same phase functions
// thread by thread to prove the indexing and the arithmetic bit-for-bit against the oracle without a GPU).
#pragma once

#ifdef __CUDACC__
#include "fast_common.cuh"
#define FHD __device__ __forceinline__
#define FHDH __host__ __device__ __forceinline__
typedef double2 d2;
#else
#include <cmath>
#include <cstring>
#define FHD static inline
#define FHDH static inline
struct d2 {
    double x, y;
};
struct DivC {
    double c, rc, rl;
};
static inline DivC divc_of(double c) {
    DivC d;
    d.c = c; d.rc = 1.0 / c; d.rl = std::fma(-c, d.rc, 1.0) / c;
    return d;
}
struct Strides {
    int sy, sz;
};
struct InclDev {
    int    active;
    int    nd;
    int    loc[3];
    double origin[3], spacing[3], c0[3];
    double r2, in, out;
};
using std::fma;
template <bool TRUE_DIV>
static inline double div_u(double x, const DivC d) {
    if (TRUE_DIV) return x / d.c;
    const double q = fma(x, d.rc, x * d.rl);
    const double r = fma(-d.c, q, x);
    return fma(r, d.rc, q);
}
template <int MODE>
static inline double div_m(double x, const DivC d) {
    if (MODE == 2) return fma(x, d.rc, x * d.rl);
    return div_u<MODE == 1>(x, d);
}
static inline double coord_dev(double origin, double spacing, int loc, int i) {
    const double im1 = (double)(i - 1);
    return loc == 1 ? fma(im1, spacing, origin) : fma(im1, spacing, fma(0.5, spacing, origin));
}
static inline d2 ld2(const double* p) { return d2{p[0], p[1]}; }
static inline void st2(double* p, d2 v) { p[0] = v.x; p[1] = v.y; }
#endif

constexpr int FSV_LANES = 32;
constexpr int FSV_XI    = 60;   // interior cells of a 64-cell row segment
constexpr int FSV_NF    = 7;    // published per plane: Pr xx yy zz xy xz yz
enum { FSV_PR = 0, FSV_XX, FSV_YY, FSV_ZZ, FSV_XY, FSV_XZ, FSV_YZ };

// Launch order of the tiles (clusters) of one sweep.  Natural order: x fastest, then y, then z -- tiles that run at the
// same time are neighbours in space and share their halo rows / lanes in L2.  With `tail` set (a launch whose boundary
// batches / halo exchange overlap the sweep, api.cu run_overlapped) the order is arranged so that every BOUNDARY tile --
// one that owns cells the batches or the exchange read or write: the first and last tiles along each dimension -- has
// retired while a last stretch of interior tiles is still running: the last z layer of tiles (all boundary) runs first,
// the other layers follow in natural order, and inside the final layer the boundary strips go before its interior.
// Measured on B200 (2 GPUs, 767^3): launching ALL boundary tiles first costs the sweep 0.3 ms of L2 locality; this order
// keeps the natural one for all but ~1.5 % of the tiles and still leaves ~6 % of the sweep to hide the exchange behind.
// g: tiles per dim; [i0, i1): interior tile indices per dim (KernelLaunch.jl:160-181 overlaps the same work by splitting
// the launch into an inner region and six slabs).
struct TileOrder {
    int g[3], i0[3], i1[3];
    int tail;
};
FHDH int tile_total(const TileOrder& o) { return o.g[0] * o.g[1] * o.g[2]; }
FHDH int tile_interior(const TileOrder& o) { return (o.i1[0] - o.i0[0]) * (o.i1[1] - o.i0[1]) * (o.i1[2] - o.i0[2]); }
// linear index c -> tile (bx, by, bz); returns true for a boundary tile
FHDH bool tile_decode(const TileOrder& o, int c, int& bx, int& by, int& bz) {
    const int layer = o.g[0] * o.g[1];
    const int ci = c / layer;
    int r = c - ci * layer;
    if (!o.tail) {
        bz = ci; bx = r % o.g[0]; by = r / o.g[0];
    } else {
        bz = ci == 0 ? o.g[2] - 1 : ci - 1;
        if (ci < o.g[2] - 1) {
            bx = r % o.g[0]; by = r / o.g[0];
        } else {        // the final layer: whole rows of the first / last y tiles, then the first / last x tiles, then the rest
            const int ix = o.i1[0] - o.i0[0], iy = o.i1[1] - o.i0[1];
            const int xb = o.g[0] - ix, yb = o.g[1] - iy, py = o.g[0] * yb;
            int t;
            if (r < py) {
                bx = r % o.g[0]; t = r / o.g[0];
                by = t < o.i0[1] ? t : t - o.i0[1] + o.i1[1];
            } else if (r - py < xb * iy) {
                r -= py;
                t = r % xb; by = o.i0[1] + r / xb;
                bx = t < o.i0[0] ? t : t - o.i0[0] + o.i1[0];
            } else {
                r -= py + xb * iy;
                bx = o.i0[0] + r % ix; by = o.i0[1] + r / ix;
            }
        }
    }
    return bx < o.i0[0] || bx >= o.i1[0] || by < o.i0[1] || by >= o.i1[1] || bz < o.i0[2] || bz >= o.i1[2];
}

struct FusedP {
    const double *tc[6], *to[6], *Prc, *Vc[3], *rho;   // current tau, tau_old, Pr, V ; rho == nullptr -> FunctionField
    double *tn[6], *Prn, *dV, *Vn[3], *r[3];          // shadow tau / Pr / V, divV, r_V     (all at logical (0,0,0))
    Strides cc, vc, cv, vv;                            // CC: xx yy zz Pr dV Vz rz rho ; VC: Vx rx xz ; CV: Vy ry yz ; VV: xy
    int lo[3], hi[3];                                  // velocity / store box, hi exclusive
    int flo[3], fhi[3];                                // index range of the op (stresses exist only inside it)
    double idx, idy, idz, eta_ve, dtau_Pr, dtau_r, nudtau;
    DivC Gdt, eta, three, eve;
    InclDev inc;
    int cz;        // planes per z-chunk
    int rows_int;  // interior rows per cluster = CL*TYB - 2
    TileOrder order;        // linear cluster index -> tile
    unsigned int* done;     // device counter: every CTA of a boundary tile adds 1 when its stores are visible (or nullptr)
};

// shared-memory exchange buffer of one CTA: [2 plane parities][FSV_NF][TYB rows][64 cells]
FHD int fsv_xoff(int tyb, int buf, int f, int row) { return ((buf * FSV_NF + f) * tyb + row) * 64; }
static inline size_t fsv_smem_bytes(int tyb) { return (size_t)2 * FSV_NF * tyb * 64 * sizeof(double); }

struct FusedT {
    int  lane, ty, i, j, k0, k1;
    bool s_act;           // this lane loads and computes stresses
    int  role;            // 0: interior row; 1: first row of the cluster (only Pr, tau_yy are consumed, by the row above);
                          // 2: last row (only tau_xy, tau_yz are consumed, by the row below)
    int  nv;              // cells of the pair that are updated and stored (0, 1 or 2)
    bool fx0, fx1, fy;    // cell / row inside the op's index range
    int  jm, jp;          // 1 if row j-1 / j+1 may be addressed (0 on the cluster's first / last row)
    long long cc, vc, cv, vv;                              // element offsets of (i, j, kp)
    d2 vx_km, vy_km, vz_km, vz_k, vzjm;                    // V planes carried along z
    d2 vx_k, vy_k, vz_kp, vzjm_kp;                         // phase A -> phase B
    d2 pr_km, tzz_km;                                      // new Pr, tau_zz of plane kp-2 (for the velocity of kp-1)
    double sxy0, sxy1;                                     // FunctionField rho_g: x,y part of the squared radius
};

FHD d2 fsv_zero() {
    d2 z;
    z.x = 0.0; z.y = 0.0;
    return z;
}

// geometry of a thread: bx = row-segment index along x, grow = row index inside the cluster (0 .. rows_int+1),
// cyc = cluster index along y, bz = z-chunk index
FHD void fsv_init(FusedT& s, const FusedP& p, int lane, int ty, int grow, int bx, int cyc, int bz, bool fun) {
    s.lane = lane; s.ty = ty;
    s.i  = p.lo[0] - 2 + bx * FSV_XI + 2 * lane;
    s.j  = p.lo[1] - 1 + cyc * p.rows_int + grow;
    s.k0 = p.lo[2] + bz * p.cz;
    s.k1 = s.k0 + p.cz < p.hi[2] ? s.k0 + p.cz : p.hi[2];
    s.s_act = s.i <= p.hi[0] && s.j <= p.hi[1];
    const bool vrow = grow >= 1 && grow <= p.rows_int && s.j < p.hi[1];
    const bool vlane = lane >= 1 && lane <= FSV_LANES - 2;
    int nv = p.hi[0] - s.i;
    nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
    s.nv = (vrow && vlane) ? nv : 0;
    s.fx0 = s.i >= p.flo[0] && s.i < p.fhi[0];
    s.fx1 = s.i + 1 >= p.flo[0] && s.i + 1 < p.fhi[0];
    s.fy  = s.j >= p.flo[1] && s.j < p.fhi[1];
    s.jm  = grow >= 1 ? 1 : 0;
    s.jp  = grow <= p.rows_int ? 1 : 0;
    s.role = grow == 0 ? 1 : (grow == p.rows_int + 1 ? 2 : 0);
    const int kp = s.k0 - 1;
    s.cc = (long long)s.i + (long long)s.j * p.cc.sy + (long long)kp * p.cc.sz;
    s.vc = (long long)s.i + (long long)s.j * p.vc.sy + (long long)kp * p.vc.sz;
    s.cv = (long long)s.i + (long long)s.j * p.cv.sy + (long long)kp * p.cv.sz;
    s.vv = (long long)s.i + (long long)s.j * p.vv.sy + (long long)kp * p.vv.sz;
    const d2 z = fsv_zero();
    s.vx_km = z; s.vy_km = z; s.vz_km = z; s.vz_k = z; s.vzjm = z;
    s.vx_k = z; s.vy_k = z; s.vz_kp = z; s.vzjm_kp = z; s.pr_km = z; s.tzz_km = z;
    if (s.s_act) {
        // plane k0-2 is never needed (only Pr and tau_zz of plane k0-1 are consumed): stand in with plane k0-1
        s.vx_km = ld2(p.Vc[0] + s.vc);
        s.vy_km = ld2(p.Vc[1] + s.cv);
        s.vz_k  = ld2(p.Vc[2] + s.cc);
        s.vzjm  = ld2(p.Vc[2] + s.cc - (long long)s.jm * p.cc.sy);
    }
    s.sxy0 = 0.0; s.sxy1 = 0.0;
    if (fun) {
        const double cy = coord_dev(p.inc.origin[1], p.inc.spacing[1], p.inc.loc[1], s.j) - p.inc.c0[1];
        const double c0 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], s.i) - p.inc.c0[0];
        const double c1 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], s.i + 1) - p.inc.c0[0];
        s.sxy0 = c0 * c0 + cy * cy;
        s.sxy1 = c1 * c1 + cy * cy;
    }
}

template <int TD>
FHD double fsv_stress_upd(double t, double to, double e2, const FusedP& p) {
    // tau + (((-(tau - tau_old))/(G dt) - tau/eta) + 2 e) * eta_ve * dtau_r      (stokes_3d_inc_ve_T.jl:34-45)
    const double r = (div_m<TD>(-(t - to), p.Gdt) - div_m<TD>(t, p.eta)) + e2;
    return t + (r * p.eta_ve) * p.dtau_r;
}

// x-neighbours of OLD velocities: register shuffles on the device, plain loads in the host emulation
#ifdef __CUDACC__
FHD double fsv_from_right(double own_x, const double*, bool) { return __shfl_down_sync(0xffffffffu, own_x, 1); }
FHD double fsv_from_left(double own_y, const double*, bool) { return __shfl_up_sync(0xffffffffu, own_y, 1); }
#else
FHD double fsv_from_right(double, const double* p_ip2, bool ok) { return ok ? *p_ip2 : 0.0; }
FHD double fsv_from_left(double, const double* p_im1, bool ok) { return ok ? *p_im1 : 0.0; }
#endif

// ---- phase A: stresses of plane kp -> sn[FSV_NF] (new Pr, tau), stores for the cells this thread owns
// A thread loads only what somebody consumes.  In y: the two halo rows of a cluster feed Pr and tau_yy to the row above
// (first row) resp. tau_xy and tau_yz to the row below (last row): 7 resp. 9 of the 19 loads (measured at 767^3: 20.31 ->
// 20.22 ms, profiles/r2_c13_tune_fused.log).  In z: of the warm-up plane k0-1 only Pr and tau_zz are consumed (by the
// z-velocity of plane k0) and of the closing plane k1 only tau_xz and tau_yz (by the velocity of plane k1-1), both by the
// thread itself; besides those a thread requests the velocities the NEXT plane takes from its carried registers.
enum {
    FSV_L_VX = 1, FSV_L_VXJM = 2, FSV_L_VY = 4, FSV_L_VYJP = 8, FSV_L_VZKP = 16, FSV_L_VZJMKP = 32, FSV_L_PR = 64,
    FSV_L_T0 = 128   // tau / tau_old component c: FSV_L_T0 << c
};
FHD int fsv_need(int role, bool first, bool last) {
    const int T = FSV_L_T0;
    if (last) return role == 0 ? (FSV_L_VX | FSV_L_VY | T << 4 | T << 5) : 0;
    if (first) {
        if (role == 0) return FSV_L_VX | FSV_L_VY | FSV_L_VYJP | FSV_L_VZKP | FSV_L_VZJMKP | FSV_L_PR | T << 2;
        return role == 1 ? FSV_L_VZKP : (FSV_L_VY | FSV_L_VZKP | FSV_L_VZJMKP);
    }
    return -1;                      // a plane of the chunk proper: everything the row's role asks for (fsv_phase_a)
}

template <int TD>
FHD void fsv_phase_a(FusedT& s, const FusedP& p, int kp, d2 sn[FSV_NF]) {
    const d2 z2 = fsv_zero();
    d2 vx = z2, vxjm = z2, vy = z2, vyjp = z2, vzkp = z2, vzjmkp = z2, pr = z2;
    d2 t[6], o[6];
    if (kp < s.k0 || kp >= s.k1) {             // warm-up / closing plane of the chunk (2 of cz + 2): request by request
#pragma unroll
        for (int c = 0; c < 6; ++c) { t[c] = z2; o[c] = z2; }
        const int need = s.s_act ? fsv_need(s.role, kp < s.k0, kp >= s.k1) : 0;
        if (need & FSV_L_VX) vx = ld2(p.Vc[0] + s.vc);
        if (need & FSV_L_VY) vy = ld2(p.Vc[1] + s.cv);
        if (need & FSV_L_VYJP) vyjp = ld2(p.Vc[1] + s.cv + (long long)s.jp * p.cv.sy);
        if (need & FSV_L_VZKP) vzkp = ld2(p.Vc[2] + s.cc + p.cc.sz);
        if (need & FSV_L_VZJMKP) vzjmkp = ld2(p.Vc[2] + s.cc - (long long)s.jm * p.cc.sy + p.cc.sz);
        if (need & FSV_L_PR) pr = ld2(p.Prc + s.cc);
        if (need & (FSV_L_T0 << 2)) { t[2] = ld2(p.tc[2] + s.cc); o[2] = ld2(p.to[2] + s.cc); }
        if (need & (FSV_L_T0 << 4)) { t[4] = ld2(p.tc[4] + s.vc); o[4] = ld2(p.to[4] + s.vc); }
        if (need & (FSV_L_T0 << 5)) { t[5] = ld2(p.tc[5] + s.cv); o[5] = ld2(p.to[5] + s.cv); }
    } else if (s.role != 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { t[c] = z2; o[c] = z2; }
        if (s.s_act && s.role == 1) {          // divV (all three normal strain rates), Pr, tau_yy
            vx   = ld2(p.Vc[0] + s.vc);
            vy   = ld2(p.Vc[1] + s.cv);
            vyjp = ld2(p.Vc[1] + s.cv + (long long)s.jp * p.cv.sy);
            vzkp = ld2(p.Vc[2] + s.cc + p.cc.sz);
            pr   = ld2(p.Prc + s.cc);
            t[1] = ld2(p.tc[1] + s.cc); o[1] = ld2(p.to[1] + s.cc);
        } else if (s.s_act) {                  // the two shear strain rates with a y index, tau_xy, tau_yz
            vx     = ld2(p.Vc[0] + s.vc);
            vxjm   = ld2(p.Vc[0] + s.vc - (long long)s.jm * p.vc.sy);
            vy     = ld2(p.Vc[1] + s.cv);
            vzkp   = ld2(p.Vc[2] + s.cc + p.cc.sz);
            vzjmkp = ld2(p.Vc[2] + s.cc - (long long)s.jm * p.cc.sy + p.cc.sz);
            t[3] = ld2(p.tc[3] + s.vv); o[3] = ld2(p.to[3] + s.vv);
            t[5] = ld2(p.tc[5] + s.cv); o[5] = ld2(p.to[5] + s.cv);
        }
    } else if (s.s_act) {
        vx     = ld2(p.Vc[0] + s.vc);
        vxjm   = ld2(p.Vc[0] + s.vc - (long long)s.jm * p.vc.sy);
        vy     = ld2(p.Vc[1] + s.cv);
        vyjp   = ld2(p.Vc[1] + s.cv + (long long)s.jp * p.cv.sy);
        vzkp   = ld2(p.Vc[2] + s.cc + p.cc.sz);
        vzjmkp = ld2(p.Vc[2] + s.cc - (long long)s.jm * p.cc.sy + p.cc.sz);
        pr = ld2(p.Prc + s.cc);
#pragma unroll
        for (int c = 0; c < 3; ++c) { t[c] = ld2(p.tc[c] + s.cc); o[c] = ld2(p.to[c] + s.cc); }
        t[3] = ld2(p.tc[3] + s.vv); o[3] = ld2(p.to[3] + s.vv);
        t[4] = ld2(p.tc[4] + s.vc); o[4] = ld2(p.to[4] + s.vc);
        t[5] = ld2(p.tc[5] + s.cv); o[5] = ld2(p.to[5] + s.cv);
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) { t[c] = z2; o[c] = z2; }
    }
    // Vx[i+2], Vy[i-1], Vz[i-1]: lanes 31 / 0 get a don't-care value (their outer cell's stresses are never used)
    const bool okr = s.s_act && s.lane < FSV_LANES - 1, okl = s.s_act && s.lane > 0;
    const double vx_ip2 = fsv_from_right(vx.x, p.Vc[0] + s.vc + 2, okr);
    const double vy_im1 = fsv_from_left(vy.y, p.Vc[1] + s.cv - 1, okl);
    const double vz_im1 = fsv_from_left(s.vz_k.y, p.Vc[2] + s.cc - 1, okl);

    const bool fz = kp >= p.flo[2] && kp < p.fhi[2];
    d2 dv = z2, prn = z2, tn[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double a_vx = h ? vx.y : vx.x, a_vxip = h ? vx_ip2 : vx.y, a_vxjm = h ? vxjm.y : vxjm.x, a_vxkm = h ? s.vx_km.y : s.vx_km.x;
        const double a_vy = h ? vy.y : vy.x, a_vyjp = h ? vyjp.y : vyjp.x, a_vyim = h ? vy.x : vy_im1, a_vykm = h ? s.vy_km.y : s.vy_km.x;
        const double a_vz = h ? s.vz_k.y : s.vz_k.x, a_vzkp = h ? vzkp.y : vzkp.x, a_vzim = h ? s.vz_k.x : vz_im1, a_vzjm = h ? s.vzjm.y : s.vzjm.x;
        const double exx = (a_vxip - a_vx) * p.idx;
        const double eyy = (a_vyjp - a_vy) * p.idy;
        const double ezz = (a_vzkp - a_vz) * p.idz;
        const double exy = 0.5 * ((a_vx - a_vxjm) * p.idy + (a_vy - a_vyim) * p.idx);
        const double exz = 0.5 * ((a_vx - a_vxkm) * p.idz + (a_vz - a_vzim) * p.idx);
        const double eyz = 0.5 * ((a_vy - a_vykm) * p.idz + (a_vz - a_vzjm) * p.idy);
        const double d   = (exx + eyy) + ezz;
        const double d3  = div_m<TD>(d, p.three);
        const double a_pr = h ? pr.y : pr.x;
        const double e2[6] = {2.0 * (exx - d3), 2.0 * (eyy - d3), 2.0 * (ezz - d3), 2.0 * exy, 2.0 * exz, 2.0 * eyz};
        // outside the op's index range update_stress! never ran: the value the velocity update sees is the stored one
        const bool in = (h ? s.fx1 : s.fx0) && s.fy && fz;
        const double n_pr = in ? a_pr - (d * p.eta_ve) * p.dtau_Pr : a_pr;
        if (h) { dv.y = d; prn.y = n_pr; } else { dv.x = d; prn.x = n_pr; }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const double tc = h ? t[c].y : t[c].x;
            const double r  = in ? fsv_stress_upd<TD>(tc, h ? o[c].y : o[c].x, e2[c], p) : tc;
            if (h) tn[c].y = r; else tn[c].x = r;
        }
    }
    if (kp >= s.k0 && kp < s.k1) {
        if (s.nv == 2) {
            st2(p.dV + s.cc, dv);
            st2(p.Prn + s.cc, prn);
#pragma unroll
            for (int c = 0; c < 3; ++c) st2(p.tn[c] + s.cc, tn[c]);
            st2(p.tn[3] + s.vv, tn[3]);
            st2(p.tn[4] + s.vc, tn[4]);
            st2(p.tn[5] + s.cv, tn[5]);
        } else if (s.nv == 1) {
            p.dV[s.cc]  = dv.x;
            p.Prn[s.cc] = prn.x;
#pragma unroll
            for (int c = 0; c < 3; ++c) p.tn[c][s.cc] = tn[c].x;
            p.tn[3][s.vv] = tn[3].x;
            p.tn[4][s.vc] = tn[4].x;
            p.tn[5][s.cv] = tn[5].x;
        }
    }
    sn[FSV_PR] = prn;
#pragma unroll
    for (int c = 0; c < 6; ++c) sn[1 + c] = tn[c];
    s.vx_k = vx; s.vy_k = vy; s.vz_kp = vzkp; s.vzjm_kp = vzjmkp;
}

// ---- phase B: publish the stresses of plane kp, update the velocity of plane kp-1, rotate the carried planes.
// own / below / above: exchange buffers (element 0 of [buf][field][row][cell]) of the CTAs holding this thread's
// row, row j-1 and row j+1; rb / ra: the row numbers of j-1 / j+1 inside those CTAs.
template <int TD, bool FUN>
FHD void fsv_phase_b(FusedT& s, const FusedP& p, int kp, const d2 sn[FSV_NF], int tyb, double* own, const double* below,
                     int rb, const double* above, int ra) {
    const int cur = kp & 1, prev = cur ^ 1;
    const int c2 = 2 * s.lane;
#pragma unroll
    for (int f = 0; f < FSV_NF; ++f) st2(own + fsv_xoff(tyb, cur, f, s.ty) + c2, sn[f]);
    if (s.nv > 0 && kp >= s.k0) {
        const d2 pr  = ld2(own + fsv_xoff(tyb, prev, FSV_PR, s.ty) + c2);
        const d2 tzz = ld2(own + fsv_xoff(tyb, prev, FSV_ZZ, s.ty) + c2);
        if (kp >= s.k0 + 1) {   // velocity of plane k = kp-1   (stokes_3d_inc_ve_T.jl:48-57)
            const d2 txx = ld2(own + fsv_xoff(tyb, prev, FSV_XX, s.ty) + c2);
            const d2 tyy = ld2(own + fsv_xoff(tyb, prev, FSV_YY, s.ty) + c2);
            const d2 txy = ld2(own + fsv_xoff(tyb, prev, FSV_XY, s.ty) + c2);
            const d2 txz = ld2(own + fsv_xoff(tyb, prev, FSV_XZ, s.ty) + c2);
            const d2 tyz = ld2(own + fsv_xoff(tyb, prev, FSV_YZ, s.ty) + c2);
            const double pr_im1  = own[fsv_xoff(tyb, prev, FSV_PR, s.ty) + c2 - 1];
            const double txx_im1 = own[fsv_xoff(tyb, prev, FSV_XX, s.ty) + c2 - 1];
            const double txy_ip2 = own[fsv_xoff(tyb, prev, FSV_XY, s.ty) + c2 + 2];
            const double txz_ip2 = own[fsv_xoff(tyb, prev, FSV_XZ, s.ty) + c2 + 2];
            const d2 prjm  = ld2(below + fsv_xoff(tyb, prev, FSV_PR, rb) + c2);
            const d2 tyyjm = ld2(below + fsv_xoff(tyb, prev, FSV_YY, rb) + c2);
            const d2 txyjp = ld2(above + fsv_xoff(tyb, prev, FSV_XY, ra) + c2);
            const d2 tyzjp = ld2(above + fsv_xoff(tyb, prev, FSV_YZ, ra) + c2);
            const d2 txzkp = sn[FSV_XZ], tyzkp = sn[FSV_YZ];
            const long long cc = s.cc - p.cc.sz, vc = s.vc - p.vc.sz, cv = s.cv - p.cv.sz;
            d2 rho;
            if (FUN) {
                const double cz  = coord_dev(p.inc.origin[2], p.inc.spacing[2], p.inc.loc[2], kp - 1) - p.inc.c0[2];
                const double cz2 = cz * cz;
                rho.x = (s.sxy0 + cz2) < p.inc.r2 ? p.inc.in : p.inc.out;
                rho.y = (s.sxy1 + cz2) < p.inc.r2 ? p.inc.in : p.inc.out;
            } else if (s.nv == 2) {
                rho = ld2(p.rho + cc);
            } else {
                rho.x = p.rho[cc]; rho.y = 0.0;
            }
            d2 nrx, nry, nrz, nvx, nvy, nvz;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double a_pr = h ? pr.y : pr.x, a_prim = h ? pr.x : pr_im1, a_prjm = h ? prjm.y : prjm.x, a_prkm = h ? s.pr_km.y : s.pr_km.x;
                const double a_txx = h ? txx.y : txx.x, a_txxim = h ? txx.x : txx_im1;
                const double a_tyy = h ? tyy.y : tyy.x, a_tyyjm = h ? tyyjm.y : tyyjm.x;
                const double a_tzz = h ? tzz.y : tzz.x, a_tzzkm = h ? s.tzz_km.y : s.tzz_km.x;
                const double a_txy = h ? txy.y : txy.x, a_txyjp = h ? txyjp.y : txyjp.x, a_txyip = h ? txy_ip2 : txy.y;
                const double a_txz = h ? txz.y : txz.x, a_txzkp = h ? txzkp.y : txzkp.x, a_txzip = h ? txz_ip2 : txz.y;
                const double a_tyz = h ? tyz.y : tyz.x, a_tyzkp = h ? tyzkp.y : tyzkp.x, a_tyzjp = h ? tyzjp.y : tyzjp.x;
                const double rvx = (((-((a_pr - a_prim) * p.idx)) + (a_txx - a_txxim) * p.idx) + (a_txyjp - a_txy) * p.idy) +
                                   (a_txzkp - a_txz) * p.idz;
                const double rvy = (((-((a_pr - a_prjm) * p.idy)) + (a_tyy - a_tyyjm) * p.idy) + (a_txyip - a_txy) * p.idx) +
                                   (a_tyzkp - a_tyz) * p.idz;
                const double rvz = ((((-((a_pr - a_prkm) * p.idz)) + (a_tzz - a_tzzkm) * p.idz) + (a_txzip - a_txz) * p.idx) +
                                    (a_tyzjp - a_tyz) * p.idy) - (h ? rho.y : rho.x);
                const double ux = (h ? s.vx_km.y : s.vx_km.x) + div_m<TD>(rvx * p.nudtau, p.eve);
                const double uy = (h ? s.vy_km.y : s.vy_km.x) + div_m<TD>(rvy * p.nudtau, p.eve);
                const double uz = (h ? s.vz_km.y : s.vz_km.x) + div_m<TD>(rvz * p.nudtau, p.eve);
                if (h) { nrx.y = rvx; nry.y = rvy; nrz.y = rvz; nvx.y = ux; nvy.y = uy; nvz.y = uz; }
                else   { nrx.x = rvx; nry.x = rvy; nrz.x = rvz; nvx.x = ux; nvy.x = uy; nvz.x = uz; }
            }
            if (s.nv == 2) {
                st2(p.r[0] + vc, nrx); st2(p.r[1] + cv, nry); st2(p.r[2] + cc, nrz);
                st2(p.Vn[0] + vc, nvx); st2(p.Vn[1] + cv, nvy); st2(p.Vn[2] + cc, nvz);
            } else {
                p.r[0][vc] = nrx.x; p.r[1][cv] = nry.x; p.r[2][cc] = nrz.x;
                p.Vn[0][vc] = nvx.x; p.Vn[1][cv] = nvy.x; p.Vn[2][cc] = nvz.x;
            }
        }
        s.pr_km = pr; s.tzz_km = tzz;
    }
    s.vx_km = s.vx_k; s.vy_km = s.vy_k; s.vz_km = s.vz_k; s.vz_k = s.vz_kp; s.vzjm = s.vzjm_kp;
    s.cc += p.cc.sz; s.vc += p.vc.sz; s.cv += p.cv.sz; s.vv += p.vv.sz;
}
