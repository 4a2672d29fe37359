// fused_sv2d.cuh -- EXPERIMENTAL (round-2 candidate; opt-in through chmy_set_fusion(ctx, 3); proven by the host
// emulation, not yet run on a GPU): update_stress! + update_velocity! of the 2D Stokes PT iteration
// (examples/stokes_2d_inc_ve_T.jl:20-43) in ONE y-marching sweep, the 2D sibling of fused_sv.cuh.
//
// Two kernels move 24 array passes per iteration (stress R9+W5, velocity R6+W4); one sweep moves R9 + W9 = 18.
// In 2D the sweep needs neither shared memory nor barriers: a lane owns 2 x-adjacent cells, a warp covers 64 cells of
// which lanes 1..30 are interior (lanes 0 / 31 recompute the stresses the interior needs from the neighbouring row
// segment), and the warp marches along y.  Iteration jp computes the stresses of row jp (phase A) and then the
// velocity of row jp-1 (phase B): x-neighbours of the NEW stresses come from warp shuffles, the row jp-1 / jp-2
// values are carried in registers, row jp is phase A's result.  tau, Pr and V are read from the current buffers and
// written to the shadow buffers (ping-pong, as in 3D); outside the op's index range the "new" value is the stored one.
//
// Compiles with nvcc (ops_fused2d.cu) and with a host compiler (tests/emul/fused_emul2d.cpp runs the 32 lanes of a
// warp in lock-step; the two shuffles of phase B are the only lines that differ).
#pragma once
#include "fused_sv.cuh"   // d2, DivC, Strides, InclDev, ld2/st2, div_u, coord_dev, fsv_from_left/right, FSV_* constants

struct Fused2P {
    const double *tc[3], *to[3], *Prc, *Vc[2], *rho;   // current tau (xx yy xy), tau_old, Pr, V ; rho == nullptr -> FunctionField
    double *tn[3], *Prn, *dV, *Vn[2], *r[2];          // shadow tau / Pr / V, divV, r_V    (all at logical (0,0))
    int s_cc, s_vc, s_cv, s_vv;                        // row strides: CC xx yy Pr dV ; VC Vx rx ; CV Vy ry rho ; VV xy
    int lo[2], hi[2];                                  // velocity / store box, hi exclusive
    int flo[2], fhi[2];                                // index range of the op
    double idx, idy, eta_ve, dtau_Pr, dtau_r, nudtau;
    DivC Gdt, eta, three, eve;
    InclDev inc;
    int cy;                                            // rows per y-chunk
};

struct Fused2T {
    int  lane, i, j0, j1;
    bool s_act;            // this lane loads and computes stresses
    int  nv;               // cells of the pair that are updated and stored (0, 1 or 2)
    bool fx0, fx1;
    long long cc, vc, cv, vv;          // element offsets of (i, jp)
    d2 vx_jm, vy_jm, vy_j;             // Vx(jp-1), Vy(jp-1), Vy(jp)
    d2 prC, txxC, tyyC, txyC;          // new values of row jp-1
    d2 pr_jm, tyy_jm;                  // new Pr, tau_yy of row jp-2
    d2 vx_j, vy_jp;                    // phase A -> rotation in phase B: Vx(jp), Vy(jp+1)
    double sx0, sx1;                   // FunctionField rho_g: x part of the squared radius
};

// seg = row-segment index along x, cyc = y-chunk index
FHD void fsv2_init(Fused2T& s, const Fused2P& p, int lane, int seg, int cyc, bool fun) {
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
    const int jp = s.j0 - 1;
    s.cc = (long long)s.i + (long long)jp * p.s_cc;
    s.vc = (long long)s.i + (long long)jp * p.s_vc;
    s.cv = (long long)s.i + (long long)jp * p.s_cv;
    s.vv = (long long)s.i + (long long)jp * p.s_vv;
    const d2 z = fsv_zero();
    s.vx_jm = s.vy_jm = s.vy_j = s.prC = s.txxC = s.tyyC = s.txyC = s.pr_jm = s.tyy_jm = s.vx_j = s.vy_jp = z;
    if (s.s_act) {
        s.vx_jm = ld2(p.Vc[0] + s.vc);       // row j0-2 is never needed (only Pr, tau_yy of row j0-1 are consumed)
        s.vy_j  = ld2(p.Vc[1] + s.cv);
    }
    s.sx0 = s.sx1 = 0.0;
    if (fun) {
        const double c0 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], s.i) - p.inc.c0[0];
        const double c1 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], s.i + 1) - p.inc.c0[0];
        s.sx0 = c0 * c0;
        s.sx1 = c1 * c1;
    }
}

template <int TD>
FHD double fsv2_stress_upd(double t, double to, double e2, const Fused2P& p) {
    const double r = (div_m<TD>(-(t - to), p.Gdt) - div_m<TD>(t, p.eta)) + e2;   // stokes_2d_inc_ve_T.jl:30-33
    return t + (r * p.eta_ve) * p.dtau_r;
}

// ---- phase A: stresses of row jp -> sn[4] = new Pr, xx, yy, xy; stores for the cells this thread owns
template <int TD>
FHD void fsv2_phase_a(Fused2T& s, const Fused2P& p, int jp, d2 sn[4]) {
    const d2 z2 = fsv_zero();
    d2 vx = z2, vyjp = z2, pr = z2, t[3], o[3];
    if (s.s_act) {
        vx   = ld2(p.Vc[0] + s.vc);
        vyjp = ld2(p.Vc[1] + s.cv + p.s_cv);
        pr   = ld2(p.Prc + s.cc);
        t[0] = ld2(p.tc[0] + s.cc); o[0] = ld2(p.to[0] + s.cc);
        t[1] = ld2(p.tc[1] + s.cc); o[1] = ld2(p.to[1] + s.cc);
        t[2] = ld2(p.tc[2] + s.vv); o[2] = ld2(p.to[2] + s.vv);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { t[c] = z2; o[c] = z2; }
    }
    const bool okr = s.s_act && s.lane < FSV_LANES - 1, okl = s.s_act && s.lane > 0;
    const double vx_ip2 = fsv_from_right(vx.x, p.Vc[0] + s.vc + 2, okr);
    const double vy_im1 = fsv_from_left(s.vy_j.y, p.Vc[1] + s.cv - 1, okl);
    const bool fy = jp >= p.flo[1] && jp < p.fhi[1];
    d2 dv = z2, prn = z2, tn[3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double a_vx = h ? vx.y : vx.x, a_vxip = h ? vx_ip2 : vx.y, a_vxjm = h ? s.vx_jm.y : s.vx_jm.x;
        const double a_vy = h ? s.vy_j.y : s.vy_j.x, a_vyjp = h ? vyjp.y : vyjp.x, a_vyim = h ? s.vy_j.x : vy_im1;
        const double exx = (a_vxip - a_vx) * p.idx;
        const double eyy = (a_vyjp - a_vy) * p.idy;
        const double exy = 0.5 * ((a_vx - a_vxjm) * p.idy + (a_vy - a_vyim) * p.idx);
        const double d   = exx + eyy;
        const double a_pr = h ? pr.y : pr.x;
        const bool in = (h ? s.fx1 : s.fx0) && fy;
        const double n_pr = in ? a_pr - (d * p.eta_ve) * p.dtau_Pr : a_pr;
        const double d3  = div_m<TD>(d, p.three);          // the 2D driver also divides by 3.0 (:28-29)
        const double e2[3] = {2.0 * (exx - d3), 2.0 * (eyy - d3), 2.0 * exy};
        if (h) { dv.y = d; prn.y = n_pr; } else { dv.x = d; prn.x = n_pr; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double tc = h ? t[c].y : t[c].x;
            const double r  = in ? fsv2_stress_upd<TD>(tc, h ? o[c].y : o[c].x, e2[c], p) : tc;
            if (h) tn[c].y = r; else tn[c].x = r;
        }
    }
    if (jp >= s.j0 && jp < s.j1) {
        if (s.nv == 2) {
            st2(p.dV + s.cc, dv); st2(p.Prn + s.cc, prn);
            st2(p.tn[0] + s.cc, tn[0]); st2(p.tn[1] + s.cc, tn[1]); st2(p.tn[2] + s.vv, tn[2]);
        } else if (s.nv == 1) {
            p.dV[s.cc] = dv.x; p.Prn[s.cc] = prn.x;
            p.tn[0][s.cc] = tn[0].x; p.tn[1][s.cc] = tn[1].x; p.tn[2][s.vv] = tn[2].x;
        }
    }
    sn[0] = prn; sn[1] = tn[0]; sn[2] = tn[1]; sn[3] = tn[2];
    s.vx_j = vx; s.vy_jp = vyjp;
}

// ---- phase B: velocity of row j = jp-1 from the carried new values of rows jp-1 / jp-2, phase A's row jp and the
// x-neighbours handed in by the caller (warp shuffles on the device: Pr[i-1], tau_xx[i-1] = the left lane's second
// cell, tau_xy[i+2] = the right lane's first cell, all of row jp-1); then rotate the carried rows.
template <int TD, bool FUN>
FHD void fsv2_phase_b(Fused2T& s, const Fused2P& p, int jp, const d2 sn[4], double pr_im1, double txx_im1, double txy_ip2) {
    if (s.nv > 0 && jp >= s.j0 + 1) {   // stokes_2d_inc_ve_T.jl:36-43
        const d2 pr = s.prC, txx = s.txxC, tyy = s.tyyC, txy = s.txyC, txyjp = sn[3];
        const long long vc = s.vc - p.s_vc, cv = s.cv - p.s_cv;
        d2 rho;
        if (FUN) {
            const double cy  = coord_dev(p.inc.origin[1], p.inc.spacing[1], p.inc.loc[1], jp - 1) - p.inc.c0[1];
            const double cy2 = cy * cy;
            rho.x = (s.sx0 + cy2) < p.inc.r2 ? p.inc.in : p.inc.out;
            rho.y = (s.sx1 + cy2) < p.inc.r2 ? p.inc.in : p.inc.out;
        } else if (s.nv == 2) {
            rho = ld2(p.rho + cv);
        } else {
            rho.x = p.rho[cv]; rho.y = 0.0;
        }
        d2 nrx, nry, nvx, nvy;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a_pr = h ? pr.y : pr.x, a_prim = h ? pr.x : pr_im1, a_prjm = h ? s.pr_jm.y : s.pr_jm.x;
            const double a_txx = h ? txx.y : txx.x, a_txxim = h ? txx.x : txx_im1;
            const double a_tyy = h ? tyy.y : tyy.x, a_tyyjm = h ? s.tyy_jm.y : s.tyy_jm.x;
            const double a_txy = h ? txy.y : txy.x, a_txyjp = h ? txyjp.y : txyjp.x, a_txyip = h ? txy_ip2 : txy.y;
            const double rvx = ((-((a_pr - a_prim) * p.idx)) + (a_txx - a_txxim) * p.idx) + (a_txyjp - a_txy) * p.idy;
            const double rvy = (((-((a_pr - a_prjm) * p.idy)) + (a_tyy - a_tyyjm) * p.idy) + (a_txyip - a_txy) * p.idx) -
                               (h ? rho.y : rho.x);
            const double ux = (h ? s.vx_jm.y : s.vx_jm.x) + div_m<TD>(rvx * p.nudtau, p.eve);
            const double uy = (h ? s.vy_jm.y : s.vy_jm.x) + div_m<TD>(rvy * p.nudtau, p.eve);
            if (h) { nrx.y = rvx; nry.y = rvy; nvx.y = ux; nvy.y = uy; }
            else   { nrx.x = rvx; nry.x = rvy; nvx.x = ux; nvy.x = uy; }
        }
        if (s.nv == 2) {
            st2(p.r[0] + vc, nrx); st2(p.r[1] + cv, nry); st2(p.Vn[0] + vc, nvx); st2(p.Vn[1] + cv, nvy);
        } else {
            p.r[0][vc] = nrx.x; p.r[1][cv] = nry.x; p.Vn[0][vc] = nvx.x; p.Vn[1][cv] = nvy.x;
        }
    }
    s.pr_jm = s.prC; s.tyy_jm = s.tyyC;
    s.prC = sn[0]; s.txxC = sn[1]; s.tyyC = sn[2]; s.txyC = sn[3];
    s.vx_jm = s.vx_j; s.vy_jm = s.vy_j; s.vy_j = s.vy_jp;
    s.cc += p.s_cc; s.vc += p.s_vc; s.cv += p.s_cv; s.vv += p.s_vv;
}
