// fused_emul2d.cpp -- TEST INFRASTRUCTURE.  Host execution of the phase functions of the 2D fused sweeps
// (chmy.jl_b200/csrc/fused_sv2d.cuh: update_stress! + update_velocity!; fused_pairs2d.cuh: compute_q! + update_C! and
// update_thermal_flux! + update_thermal!): the 32 lanes of a warp run in lock-step, the shuffles of the kernels are
// reads of the neighbouring lane's state taken before anybody rotates it.  tests/test_fused_emulation2d.py compares
// the results bit-for-bit with the oracle's two ops run one after the other.
#include <vector>

#include "../../chmy.jl_b200/csrc/fused_sv2d.cuh"
#include "../../chmy.jl_b200/csrc/fused_pairs2d.cuh"

template <int TD, bool FUN>
static void run2(const Fused2P& p) {
    const int nx = p.hi[0] - p.lo[0], ny = p.hi[1] - p.lo[1];
    const int gx = (nx + FSV_XI - 1) / FSV_XI, gy = (ny + p.cy - 1) / p.cy;
    Fused2T T[FSV_LANES];
    d2 SN[FSV_LANES][4];
    double L1[FSV_LANES], L2[FSV_LANES], R1[FSV_LANES];
    for (int cyc = 0; cyc < gy; ++cyc)
        for (int seg = 0; seg < gx; ++seg) {
            for (int lane = 0; lane < FSV_LANES; ++lane) fsv2_init(T[lane], p, lane, seg, cyc, FUN);
            for (int jp = T[0].j0 - 1; jp <= T[0].j1; ++jp) {
                for (int lane = 0; lane < FSV_LANES; ++lane) fsv2_phase_a<TD>(T[lane], p, jp, SN[lane]);
                for (int lane = 0; lane < FSV_LANES; ++lane) {      // __shfl_up / __shfl_down by one lane
                    L1[lane] = lane > 0 ? T[lane - 1].prC.y : T[lane].prC.y;
                    L2[lane] = lane > 0 ? T[lane - 1].txxC.y : T[lane].txxC.y;
                    R1[lane] = lane < FSV_LANES - 1 ? T[lane + 1].txyC.x : T[lane].txyC.x;
                }
                for (int lane = 0; lane < FSV_LANES; ++lane) fsv2_phase_b<TD, FUN>(T[lane], p, jp, SN[lane], L1[lane], L2[lane], R1[lane]);
            }
        }
}

// ptrs: tc[3] to[3] Prc Vc[2] rho tn[3] Prn dV Vn[2] r[2]  (19 pointers at logical (0,0); rho may be NULL)
// strides: s_cc s_vc s_cv s_vv ; box: lo[2] hi[2] flo[2] fhi[2]
// sc: idx idy eta_ve dtau_Pr dtau_r nudtau Gdt eta ; inc: origin[2] spacing[2] c0[2] r2 in out
extern "C" int fused_emul2d_run(double** ptrs, const int* strides, const int* box, const double* sc, const double* inc,
                                const int* incloc, int cy, int td) {
    Fused2P p;
    memset(&p, 0, sizeof(p));
    int q = 0;
    for (int c = 0; c < 3; ++c) p.tc[c] = ptrs[q++];
    for (int c = 0; c < 3; ++c) p.to[c] = ptrs[q++];
    p.Prc = ptrs[q++];
    for (int c = 0; c < 2; ++c) p.Vc[c] = ptrs[q++];
    p.rho = ptrs[q++];
    for (int c = 0; c < 3; ++c) p.tn[c] = ptrs[q++];
    p.Prn = ptrs[q++];
    p.dV  = ptrs[q++];
    for (int c = 0; c < 2; ++c) p.Vn[c] = ptrs[q++];
    for (int c = 0; c < 2; ++c) p.r[c] = ptrs[q++];
    p.s_cc = strides[0]; p.s_vc = strides[1]; p.s_cv = strides[2]; p.s_vv = strides[3];
    for (int a = 0; a < 2; ++a) { p.lo[a] = box[a]; p.hi[a] = box[2 + a]; p.flo[a] = box[4 + a]; p.fhi[a] = box[6 + a]; }
    p.idx = sc[0]; p.idy = sc[1]; p.eta_ve = sc[2]; p.dtau_Pr = sc[3]; p.dtau_r = sc[4]; p.nudtau = sc[5];
    p.Gdt = divc_of(sc[6]); p.eta = divc_of(sc[7]); p.three = divc_of(3.0);
    p.eve = divc_of(sc[2]);
    p.inc.active = p.rho == nullptr; p.inc.nd = 2;
    for (int a = 0; a < 2; ++a) {
        p.inc.origin[a] = inc[a]; p.inc.spacing[a] = inc[2 + a]; p.inc.c0[a] = inc[4 + a]; p.inc.loc[a] = incloc[a];
    }
    p.inc.r2 = inc[6]; p.inc.in = inc[7]; p.inc.out = inc[8];
    p.cy = cy;
    if (p.lo[0] & 1) return -1;
    const bool fun = p.rho == nullptr;
    // td: division mode -- 0 four operations, 1 true division, 2 two operations (fast_common.cuh: div_m)
    if (td == 1)      { if (fun) run2<1, true>(p); else run2<1, false>(p); }
    else if (td == 2) { if (fun) run2<2, true>(p); else run2<2, false>(p); }
    else              { if (fun) run2<0, true>(p); else run2<0, false>(p); }
    return 0;
}

// ------------------------------------------------------------------------------------------------ flux -> update pairs
template <int KIND>
static void runq(const FusedQ2P& p, int unroll) {
    const int nx = p.hi[0] - p.lo[0], ny = p.hi[1] - p.lo[1];
    const int gx = (nx + FSV_XI - 1) / FSV_XI, gy = (ny + p.cy - 1) / p.cy;
    FusedQ2T T[FSV_LANES];
    FusedQ2L LD[FSV_LANES][8];
    d2 SN[FSV_LANES][2];
    double R1[FSV_LANES];
    for (int cyc = 0; cyc < gy; ++cyc)
        for (int seg = 0; seg < gx; ++seg) {
            for (int lane = 0; lane < FSV_LANES; ++lane) fq2_init(T[lane], p, lane, seg, cyc);
            for (int jg = T[0].j0; jg <= T[0].j1; jg += unroll) {          // the kernel's loop: groups of `unroll` rows
                for (int lane = 0; lane < FSV_LANES; ++lane)
                    for (int r = 0; r < unroll; ++r)
                        if (jg + r <= T[0].j1) fq2_load<KIND>(T[lane], p, r, LD[lane][r]);
                for (int r = 0; r < unroll && jg + r <= T[0].j1; ++r) {
                    const int jp = jg + r;
                    for (int lane = 0; lane < FSV_LANES; ++lane) fq2_phase_a<KIND>(T[lane], p, jp, LD[lane][r], SN[lane]);
                    for (int lane = 0; lane < FSV_LANES; ++lane)        // __shfl_down by one lane
                        R1[lane] = lane < FSV_LANES - 1 ? T[lane + 1].qxC.x : T[lane].qxC.x;
                    for (int lane = 0; lane < FSV_LANES; ++lane) fq2_phase_b<KIND>(T[lane], p, jp, SN[lane], R1[lane]);
                }
            }
        }
}

// ptrs: Cc Cn base qx qy Vx Vy (7 pointers at logical (0,0); base / Vx / Vy NULL for kind 0)
// strides: s_cc s_vc s_cv ; box: lo[2] hi[2] flo[2] fhi[2] ; sc: idx idy coef dt
extern "C" int fused_emul2d_pair_run(int kind, double** ptrs, const int* strides, const int* box, const double* sc, int cy,
                                     int unroll) {
    FusedQ2P p;
    memset(&p, 0, sizeof(p));
    p.Cc = ptrs[0]; p.Cn = ptrs[1]; p.base = ptrs[2]; p.qx = ptrs[3]; p.qy = ptrs[4]; p.Vx = ptrs[5]; p.Vy = ptrs[6];
    p.s_cc = strides[0]; p.s_vc = strides[1]; p.s_cv = strides[2];
    for (int a = 0; a < 2; ++a) { p.lo[a] = box[a]; p.hi[a] = box[2 + a]; p.flo[a] = box[4 + a]; p.fhi[a] = box[6 + a]; }
    p.idx = sc[0]; p.idy = sc[1]; p.coef = sc[2]; p.dt = sc[3];
    p.cy = cy;
    if (p.lo[0] & 1) return -1;
    if (unroll < 1 || unroll > 8) return -2;
    if (kind == 0) runq<0>(p, unroll); else runq<1>(p, unroll);
    return 0;
}
