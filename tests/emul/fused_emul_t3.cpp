// fused_emul_t3.cpp -- TEST INFRASTRUCTURE.  Executes the functions of the fused 3D thermal sweep
// (chmy.jl_b200/csrc/fused_thermal3.cuh, the same source nvcc compiles into k_fused_t3) on the host: the 32 lanes of a
// warp run in lock-step, the three warp shuffles of the kernel are reads of the neighbouring lanes' states.
// tests/test_fused_emulation_t3.py compares the result bit-for-bit with the oracle's update_thermal_flux! on the whole
// index range followed by update_thermal! on the box.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "../../chmy.jl_b200/csrc/fused_thermal3.cuh"

static void run(const FusedT3P& p, int rows_per_cta) {
    const int nx = p.hi[0] - p.lo[0], ny = p.hi[1] - p.lo[1], nz = p.hi[2] - p.lo[2];
    const int gx = (nx + 2 * FSV_LANES - 1) / (2 * FSV_LANES), gy = (ny + rows_per_cta - 1) / rows_per_cta, gz = (nz + p.cz - 1) / p.cz;
    FusedT3T T[FSV_LANES];
    FusedT3L L[FSV_LANES];
    double tl[FSV_LANES], tr[FSV_LANES], vr[FSV_LANES];
    for (int bz = 0; bz < gz; ++bz)
        for (int by = 0; by < gy; ++by)
            for (int ty = 0; ty < rows_per_cta; ++ty)          // one warp per row of the CTA
                for (int bx = 0; bx < gx; ++bx) {
                    for (int lane = 0; lane < FSV_LANES; ++lane) ft3_init(T[lane], p, lane, bx, by * rows_per_cta + ty, bz);
                    for (int k = T[0].k0; k < T[0].k1; ++k) {
                        for (int lane = 0; lane < FSV_LANES; ++lane) ft3_load(T[lane], p, k, L[lane]);
                        for (int lane = 0; lane < FSV_LANES; ++lane) {       // __shfl_up / __shfl_down by one lane
                            tl[lane] = lane > 0 ? T[lane - 1].t_k.y : T[lane].t_k.y;
                            tr[lane] = lane < FSV_LANES - 1 ? T[lane + 1].t_k.x : T[lane].t_k.x;
                            vr[lane] = lane < FSV_LANES - 1 ? L[lane + 1].vx.x : L[lane].vx.x;
                        }
                        for (int lane = 0; lane < FSV_LANES; ++lane) ft3_compute(T[lane], p, k, L[lane], tl[lane], tr[lane], vr[lane]);
                    }
                }
}

// ptrs: Tc Tn To qx qy qz Vx Vy Vz (logical (0,0,0)); strides: cc.sy cc.sz vc.sy vc.sz cv.sy cv.sz;
// box: lo[3] hi[3] flo[3] fhi[3]; sc: lam dt idx idy idz
extern "C" int fused_emul_t3_run(double** ptrs, const int* strides, const int* box, const double* sc, int cz, int rows_per_cta) {
    FusedT3P p;
    memset(&p, 0, sizeof(p));
    p.Tc = ptrs[0]; p.Tn = ptrs[1]; p.To = ptrs[2]; p.qx = ptrs[3]; p.qy = ptrs[4]; p.qz = ptrs[5];
    p.Vx = ptrs[6]; p.Vy = ptrs[7]; p.Vz = ptrs[8];
    p.cc = Strides{strides[0], strides[1]}; p.vc = Strides{strides[2], strides[3]}; p.cv = Strides{strides[4], strides[5]};
    for (int a = 0; a < 3; ++a) { p.lo[a] = box[a]; p.hi[a] = box[3 + a]; p.flo[a] = box[6 + a]; p.fhi[a] = box[9 + a]; }
    p.lam = sc[0]; p.dt = sc[1]; p.idx = sc[2]; p.idy = sc[3]; p.idz = sc[4];
    p.cz = cz;
    if (p.lo[0] & 1) return -1;
    run(p, rows_per_cta);
    return 0;
}
