// fused_thermal3.cuh -- EXPERIMENTAL (round-2 candidate; opt-in through chmy_set_fusion(ctx, 3); proven by the host
// emulation, not yet run on a GPU): update_thermal_flux! + update_thermal! of the 3D Stokes(+T) driver
// (examples/stokes_3d_inc_ve_T.jl:59-77, launched back to back at :167-168) as ONE z-marching sweep.
//
// Why: two kernels move R4+W3 (T, V.x, V.y, V.z -> qT.x, qT.y, qT.z) and R4+W1 (T_old, qT.* -> T) = 12 array passes per
// thermal sub-step; the second kernel only differences what the first one has just written.  One sweep that keeps the
// fluxes in registers moves R5 (T, V.*, T_old) + W4 (qT.*, T) = 9.  qT is still written: it is an output of the
// reference's first kernel and a caller may read it.
//
// How: the tiling of the tuned kernels (ops_fast.cu: a lane owns 2 x-adjacent cells, a warp one 64-cell row segment, a
// CTA TY rows, a z-chunk per CTA) with NO redundant lanes and no shared memory.  A flux is cheap (2 mul, 2 fma-free
// adds, a max and a min), so the fluxes a cell needs from its +x / +y / +z neighbours are RECOMPUTED from T and V
// instead of exchanged:
//     q.x[i+2]  from the right lane's T and V.x (two warp shuffles; the last lane of a segment loads them),
//     q.y[j+1]  from the T and V.y rows of j+1 (L1/L2 hits: the neighbouring warp loads the same rows as its own),
//     q.z[k+1]  from the T plane k+1 the march loads anyway and V.z[k+1]; it becomes q.z[k] of the next iteration.
// Recomputation is deterministic, so every flux has the bits the reference's first kernel stores.  Outside the op's
// index range [0, n+1]^3 that kernel never ran: there the update sees the STORED flux (q.x[n+2] etc.), exactly as the
// reference's second kernel does.  T is read at neighbouring cells while other threads update it, so it is read from
// the field's current buffer and written to its shadow buffer (ping-pong, swapped by the host; cells outside the op's
// range are carried over by the frame copy of ops_fused.cu).  The fluxes are written in place: nobody reads an old one
// inside the range.
//
// Arithmetic order is that of ops.cu / ops_fast.cu / the reference.  Compiles with nvcc (ops_fused2d.cu) and with a host
// compiler (tests/emul/fused_emul_t3.cpp runs the 32 lanes of a warp in lock-step; the three shuffles are the only
// lines that differ).
#pragma once
#include "fused_pairs2d.cuh"   // d2, ld2/st2, FHD, jl_max0/jl_min0 on the host, FSV_LANES

struct FusedT3P {
    const double* Tc;             // current T
    double*       Tn;             // its shadow buffer
    const double* To;             // T_old
    double *      qx, *qy, *qz;   // fluxes, written in place
    const double *Vx, *Vy, *Vz;
    Strides cc, vc, cv;           // CC: T T_old V.z q.z ; VC: V.x q.x ; CV: V.y q.y
    int lo[3], hi[3];             // update / store box, hi exclusive
    int flo[3], fhi[3];           // index range of the op
    double lam, dt, idx, idy, idz;
    int cz;                       // planes per z-chunk
};

struct FusedT3T {
    int  lane, i, j, k0, k1;
    int  nact;                    // cells of the pair inside the box (0, 1 or 2)
    bool xlast;                   // no right lane to shuffle from: T[i+2], V.x[i+2] are loaded
    bool qx1_in, qx2_in, qyjp_in; // q.x[i+1] / q.x[i+2] / q.y[j+1] lie inside the op's range (else: the stored flux)
    long long cc, vc, cv;         // element offsets of (i, j, k)
    d2 t_km, t_k, qz_k;           // carried planes
};

// what one thread requests per plane (issued before any arithmetic of the plane)
struct FusedT3L {
    d2 t_kp, t_jm, t_jp, vx, vy, vy_jp, vz_kp, to, s_qyjp, s_qzkp;
    double e_tim1, e_tip2, e_vxip2, s_qx1, s_qx2;
};

FHD double ft3_flux(double nlam, double t, double tm, double v, double id) {
    return (nlam * ((t - tm) * id) + jl_max0(v) * tm) + jl_min0(v) * t;      // stokes_3d_inc_ve_T.jl:62-70
}

FHD d2 ft3_zero() { d2 z; z.x = 0.0; z.y = 0.0; return z; }

// bx = row-segment index along x, ty/by = row inside / index of the CTA along y, bz = z-chunk index
FHD void ft3_init(FusedT3T& s, const FusedT3P& p, int lane, int bx, int row, int bz) {
    s.lane = lane;
    s.i  = p.lo[0] + (bx * FSV_LANES + lane) * 2;
    s.j  = p.lo[1] + row;
    s.k0 = p.lo[2] + bz * p.cz;
    s.k1 = s.k0 + p.cz < p.hi[2] ? s.k0 + p.cz : p.hi[2];
    int n = s.j < p.hi[1] ? p.hi[0] - s.i : 0;
    s.nact = n < 0 ? 0 : (n > 2 ? 2 : n);
    s.xlast = lane == FSV_LANES - 1 || s.i + 2 >= p.hi[0];
    s.qx1_in  = s.i + 1 < p.fhi[0];      // false only for a pair whose second cell sticks out of the range (nact == 1)
    s.qx2_in  = s.i + 2 < p.fhi[0];
    s.qyjp_in = s.j + 1 < p.fhi[1];
    s.cc = (long long)s.i + (long long)s.j * p.cc.sy + (long long)s.k0 * p.cc.sz;
    s.vc = (long long)s.i + (long long)s.j * p.vc.sy + (long long)s.k0 * p.vc.sz;
    s.cv = (long long)s.i + (long long)s.j * p.cv.sy + (long long)s.k0 * p.cv.sz;
    s.t_km = s.t_k = s.qz_k = ft3_zero();
    if (s.nact > 0 && s.k0 < s.k1) {
        s.t_km = ld2(p.Tc + s.cc - p.cc.sz);
        s.t_k  = ld2(p.Tc + s.cc);
        const d2 vz = ld2(p.Vz + s.cc);
        const double nlam = -p.lam;
        s.qz_k.x = ft3_flux(nlam, s.t_k.x, s.t_km.x, vz.x, p.idz);       // k0 >= flo[2]: always inside the range
        s.qz_k.y = ft3_flux(nlam, s.t_k.y, s.t_km.y, vz.y, p.idz);
    }
}

// plane k: every global load of the iteration
FHD void ft3_load(const FusedT3T& s, const FusedT3P& p, int k, FusedT3L& L) {
    const d2 z2 = ft3_zero();
    L.t_kp = L.t_jm = L.t_jp = L.vx = L.vy = L.vy_jp = L.vz_kp = L.to = L.s_qyjp = L.s_qzkp = z2;
    L.e_tim1 = L.e_tip2 = L.e_vxip2 = L.s_qx1 = L.s_qx2 = 0.0;
    if (s.nact == 0) return;
    L.t_kp  = ld2(p.Tc + s.cc + p.cc.sz);
    L.t_jm  = ld2(p.Tc + s.cc - p.cc.sy);
    L.t_jp  = ld2(p.Tc + s.cc + p.cc.sy);
    L.vx    = ld2(p.Vx + s.vc);
    L.vy    = ld2(p.Vy + s.cv);
    L.vy_jp = ld2(p.Vy + s.cv + p.cv.sy);
    L.vz_kp = ld2(p.Vz + s.cc + p.cc.sz);
    L.to    = ld2(p.To + s.cc);
    if (s.lane == 0) L.e_tim1 = p.Tc[s.cc - 1];
    if (s.xlast) {
        L.e_tip2  = p.Tc[s.cc + 2];
        L.e_vxip2 = p.Vx[s.vc + 2];
    }
    // outside the op's index range the flux kernel never ran: the update differences the stored flux
    if (!s.qx1_in) L.s_qx1 = p.qx[s.vc + 1];
    if (!s.qx2_in) L.s_qx2 = p.qx[s.vc + 2];
    if (!s.qyjp_in) L.s_qyjp = ld2(p.qy + s.cv + p.cv.sy);
    if (k + 1 >= p.fhi[2]) L.s_qzkp = ld2(p.qz + s.cc + p.cc.sz);
}

// plane k: fluxes of the thread's two cells (stored), the recomputed neighbour fluxes, the new T (stored to the shadow).
// sh_*: the neighbouring lanes' registers (warp shuffles on the device): T of the left lane's second cell, T and V.x of
// the right lane's first cell.
FHD void ft3_compute(FusedT3T& s, const FusedT3P& p, int k, const FusedT3L& L, double sh_t_left, double sh_t_right,
                     double sh_vx_right) {
    if (s.nact > 0) {
        const double nlam = -p.lam;
        const d2 t = s.t_k;
        const double t_im1 = s.lane == 0 ? L.e_tim1 : sh_t_left;
        const double t_ip2 = s.xlast ? L.e_tip2 : sh_t_right;
        const double vx_ip2 = s.xlast ? L.e_vxip2 : sh_vx_right;
        d2 qx, qy, qyjp, qzkp;
        qx.x = ft3_flux(nlam, t.x, t_im1, L.vx.x, p.idx);
        qx.y = s.qx1_in ? ft3_flux(nlam, t.y, t.x, L.vx.y, p.idx) : L.s_qx1;
        const double qx2 = s.qx2_in ? ft3_flux(nlam, t_ip2, t.y, vx_ip2, p.idx) : L.s_qx2;
        qy.x = ft3_flux(nlam, t.x, L.t_jm.x, L.vy.x, p.idy);
        qy.y = ft3_flux(nlam, t.y, L.t_jm.y, L.vy.y, p.idy);
        if (s.qyjp_in) {
            qyjp.x = ft3_flux(nlam, L.t_jp.x, t.x, L.vy_jp.x, p.idy);
            qyjp.y = ft3_flux(nlam, L.t_jp.y, t.y, L.vy_jp.y, p.idy);
        } else {
            qyjp = L.s_qyjp;
        }
        if (k + 1 < p.fhi[2]) {
            qzkp.x = ft3_flux(nlam, L.t_kp.x, t.x, L.vz_kp.x, p.idz);
            qzkp.y = ft3_flux(nlam, L.t_kp.y, t.y, L.vz_kp.y, p.idz);
        } else {
            qzkp = L.s_qzkp;
        }
        // the second cell of a pair that sticks out of the box along x: its q.x[i+2] operand is meaningless, and so is
        // the result -- it is not stored (nact == 1)
        d2 o;       // stokes_3d_inc_ve_T.jl:75: T = T_old - dt * divg(qT), divg folds left (field_operators.jl:50-55)
        o.x = L.to.x - p.dt * (((qx.y - qx.x) * p.idx + (qyjp.x - qy.x) * p.idy) + (qzkp.x - s.qz_k.x) * p.idz);
        o.y = L.to.y - p.dt * (((qx2 - qx.y) * p.idx + (qyjp.y - qy.y) * p.idy) + (qzkp.y - s.qz_k.y) * p.idz);
        if (s.nact == 2) {
            st2(p.qx + s.vc, qx); st2(p.qy + s.cv, qy); st2(p.qz + s.cc, s.qz_k); st2(p.Tn + s.cc, o);
        } else {
            p.qx[s.vc] = qx.x; p.qy[s.cv] = qy.x; p.qz[s.cc] = s.qz_k.x; p.Tn[s.cc] = o.x;
        }
        s.t_km = t; s.t_k = L.t_kp; s.qz_k = qzkp;
    }
    s.cc += p.cc.sz; s.vc += p.vc.sz; s.cv += p.cv.sz;
}
