// fused_emul.cpp -- TEST INFRASTRUCTURE.  Executes the phase functions of the fused stress+velocity kernel
// (chmy.jl_b200/csrc/fused_sv.cuh, the same source nvcc compiles into libchmy_b200.so) thread by thread on the host:
// every cluster of CTAs is a set of FusedT states, the shared-memory exchange buffers are host arrays (poisoned
// with NaN so that a read of a never-written slot shows up), and the barrier of the kernel is the boundary between
// the two loops over threads.  tests/test_fused_emulation.py compares the result bit-for-bit with the oracle's
// update_stress! followed by update_velocity!.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <limits>
#include <vector>

#include "../../chmy.jl_b200/csrc/fused_sv.cuh"

template <int TD, bool FUN>
static void run(const FusedP& p, int tyb, int cl) {
    const int nx = p.hi[0] - p.lo[0], ny = p.hi[1] - p.lo[1], nz = p.hi[2] - p.lo[2];
    const int gx = (nx + FSV_XI - 1) / FSV_XI, gyc = (ny + p.rows_int - 1) / p.rows_int, gz = (nz + p.cz - 1) / p.cz;
    const int nthr = cl * tyb * FSV_LANES;
    const size_t per_cta = fsv_smem_bytes(tyb) / sizeof(double);
    std::vector<FusedT> T(nthr);
    std::vector<d2> SN((size_t)nthr * FSV_NF);
    std::vector<double> smem(per_cta * cl);
    for (int bz = 0; bz < gz; ++bz)
        for (int cyc = 0; cyc < gyc; ++cyc)
            for (int bx = 0; bx < gx; ++bx) {
                for (auto& v : smem) v = std::numeric_limits<double>::quiet_NaN();
                for (int cr = 0; cr < cl; ++cr)
                    for (int ty = 0; ty < tyb; ++ty)
                        for (int lane = 0; lane < FSV_LANES; ++lane)
                            fsv_init(T[(cr * tyb + ty) * FSV_LANES + lane], p, lane, ty, cr * tyb + ty, bx, cyc, bz, FUN);
                const int k0 = T[0].k0, k1 = T[0].k1;
                for (int kp = k0 - 1; kp <= k1; ++kp) {
                    for (int t = 0; t < nthr; ++t) fsv_phase_a<TD>(T[t], p, kp, &SN[(size_t)t * FSV_NF]);
                    // ---- barrier ----
                    for (int cr = 0; cr < cl; ++cr)
                        for (int ty = 0; ty < tyb; ++ty) {
                            double* own = &smem[per_cta * cr];
                            const double *below = own, *above = own;
                            int rb = ty, ra = ty;
                            if (ty > 0) rb = ty - 1;
                            else if (cr > 0) { below = &smem[per_cta * (cr - 1)]; rb = tyb - 1; }
                            if (ty < tyb - 1) ra = ty + 1;
                            else if (cr < cl - 1) { above = &smem[per_cta * (cr + 1)]; ra = 0; }
                            for (int lane = 0; lane < FSV_LANES; ++lane) {
                                const int t = (cr * tyb + ty) * FSV_LANES + lane;
                                fsv_phase_b<TD, FUN>(T[t], p, kp, &SN[(size_t)t * FSV_NF], tyb, own, below, rb, above, ra);
                            }
                        }
                }
            }
}

// ptrs: tc[6] to[6] Prc Vc[3] rho tn[6] Prn dV Vn[3] r[3]  (31 pointers at logical (0,0,0); rho may be NULL)
// strides: cc.sy cc.sz vc.sy vc.sz cv.sy cv.sz vv.sy vv.sz ; box: lo[3] hi[3] flo[3] fhi[3]
// sc: idx idy idz eta_ve dtau_Pr dtau_r nudtau Gdt eta ; inc: origin[3] spacing[3] c0[3] r2 in out
extern "C" int fused_emul_run(double** ptrs, const int* strides, const int* box, const double* sc, const double* inc,
                              const int* incloc, int cz, int tyb, int cl, int td) {
    FusedP p;
    memset(&p, 0, sizeof(p));
    int q = 0;
    for (int c = 0; c < 6; ++c) p.tc[c] = ptrs[q++];
    for (int c = 0; c < 6; ++c) p.to[c] = ptrs[q++];
    p.Prc = ptrs[q++];
    for (int c = 0; c < 3; ++c) p.Vc[c] = ptrs[q++];
    p.rho = ptrs[q++];
    for (int c = 0; c < 6; ++c) p.tn[c] = ptrs[q++];
    p.Prn = ptrs[q++];
    p.dV  = ptrs[q++];
    for (int c = 0; c < 3; ++c) p.Vn[c] = ptrs[q++];
    for (int c = 0; c < 3; ++c) p.r[c] = ptrs[q++];
    p.cc = Strides{strides[0], strides[1]}; p.vc = Strides{strides[2], strides[3]};
    p.cv = Strides{strides[4], strides[5]}; p.vv = Strides{strides[6], strides[7]};
    for (int a = 0; a < 3; ++a) { p.lo[a] = box[a]; p.hi[a] = box[3 + a]; p.flo[a] = box[6 + a]; p.fhi[a] = box[9 + a]; }
    p.idx = sc[0]; p.idy = sc[1]; p.idz = sc[2]; p.eta_ve = sc[3]; p.dtau_Pr = sc[4]; p.dtau_r = sc[5]; p.nudtau = sc[6];
    p.Gdt = divc_of(sc[7]); p.eta = divc_of(sc[8]); p.three = divc_of(3.0);
    p.eve = divc_of(sc[3]);
    p.inc.active = p.rho == nullptr; p.inc.nd = 3;
    for (int a = 0; a < 3; ++a) {
        p.inc.origin[a] = inc[a]; p.inc.spacing[a] = inc[3 + a]; p.inc.c0[a] = inc[6 + a]; p.inc.loc[a] = incloc[a];
    }
    p.inc.r2 = inc[9]; p.inc.in = inc[10]; p.inc.out = inc[11];
    p.cz = cz; p.rows_int = cl * tyb - 2;
    if (p.lo[0] & 1) return -1;
    const bool fun = p.rho == nullptr;
    // td: division mode of the sweep -- 0 four operations, 1 true division, 2 two operations (fast_common.cuh: div_m)
    if (td == 1)      { if (fun) run<1, true>(p, tyb, cl); else run<1, false>(p, tyb, cl); }
    else if (td == 2) { if (fun) run<2, true>(p, tyb, cl); else run<2, false>(p, tyb, cl); }
    else              { if (fun) run<0, true>(p, tyb, cl); else run<0, false>(p, tyb, cl); }
    return 0;
}
