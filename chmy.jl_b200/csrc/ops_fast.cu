// ops_fast.cu -- tuned kernels for the headline configurations (selected by chmy_run_op_fast; the generic
// one-thread-per-cell kernels in ops.cu remain the fallback for every other shape or layout).
//
// 3D Stokes `update_stress!` / `update_velocity!` (examples/stokes_3d_inc_ve_T.jl:23-57), HBM-bound stencils:
//   * pitched fields: logical x-index 0 of every row is 128-byte aligned, so a warp that owns 64 consecutive
//     cells (2 per lane) moves every array with aligned 128-bit loads/stores (4 full 128 B lines per request);
//   * register-blocked z-marching: a thread walks a z-chunk of its (x-pair, y) column and keeps the k-1 / k+1
//     planes of the stencil operands in registers, so each array element is requested from DRAM once;
//   * x-neighbours come from warp shuffles (the two edge lanes issue one predicated 8-byte load), y-neighbours from
//     L1 (the neighbouring row of the same CTA requests the same lines in the same step);
//   * divisions by kernel-uniform scalars (G*dt, eta, 3.0, eta_ve) use a double-double reciprocal product plus one
//     Markstein correction (div_u, fast_common.cuh: 4 operations): bit-identical to IEEE division for every
//     normal-range operand at ~1/10 of the instruction count of the generic div.rn.f64 subroutine.  Divisors whose
//     significand is all ones, or that are not normal numbers, take the true-division instantiation.
//
// Arithmetic order is exactly that of ops.cu / the reference (compiled with -fmad=false; fma() only where written).
#include "fast_common.cuh"

// self-test kernel: counts operands for which the sequence differs from div.rn.f64 (bitwise, NaN == NaN)
template <int SEQ>      // 0: four operations (div_u), 2: two operations (div_m<2>)
__global__ void k_divcheck(DivC d, unsigned long long seed, long long n, int mode, unsigned long long* bad) {
    unsigned long long cnt = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(t + 1);   // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        double x;
        if (mode == 0) {   // random significand and sign, exponent within +-60 binades of 1
            const unsigned long long e = 1023ull - 60ull + (z >> 52) % 121ull;
            x = __longlong_as_double((long long)((z & 0x800FFFFFFFFFFFFFull) | (e << 52)));
        } else {           // integer multiples of c and their neighbours (exact and near-tie quotients)
            const double m = (double)(long long)(z >> 12);
            x = m * d.c;
            if (z & 1) x = __longlong_as_double(__double_as_longlong(x) + (long long)((z >> 1) & 3) - 1);
        }
        const double a = x / d.c, b = div_m<SEQ>(x, d);
        if (__double_as_longlong(a) != __double_as_longlong(b) && !(a != a && b != b)) ++cnt;
    }
    if (cnt) atomicAdd(bad, cnt);
}

extern "C" int chmy_selftest_division(chmy_ctx* ctx, double c, long long n, unsigned long long seed,
                                      unsigned long long* mismatches, int* markstein_used) {
    CHMY_REQUIRE(ctx && mismatches && markstein_used && n >= 0, "bad argument");
    *markstein_used = markstein_ok(c) ? 1 : 0;
    *mismatches = 0;
    if (!*markstein_used) return CHMY_OK;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaMemsetAsync(ctx->d_red, 0, sizeof(unsigned long long), ctx->s_main));
    const DivC d = divc_of(c);
    k_divcheck<0><<<ctx->sm_count * 8, 256, 0, ctx->s_main>>>(d, seed, n / 2, 0, ctx->d_red);
    k_divcheck<0><<<ctx->sm_count * 8, 256, 0, ctx->s_main>>>(d, seed ^ 0xABCDEFull, n - n / 2, 1, ctx->d_red);
    ctx->n_launches += 2;
    CHMY_CUDA(cudaGetLastError());
    CHMY_CUDA(cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    *mismatches = ctx->h_red[0];
    return CHMY_OK;
}

// the two-operation sequence on the device, for a divisor div2_exact() has proven (proved = 0: nothing is run)
extern "C" int chmy_selftest_division2(chmy_ctx* ctx, double c, long long n, unsigned long long seed,
                                       unsigned long long* mismatches, int* proved) {
    CHMY_REQUIRE(ctx && mismatches && proved && n >= 0, "bad argument");
    *proved = div2_exact(c) ? 1 : 0;
    *mismatches = 0;
    if (!*proved) return CHMY_OK;
    CHMY_CUDA(cudaSetDevice(ctx->device));
    CHMY_CUDA(cudaMemsetAsync(ctx->d_red, 0, sizeof(unsigned long long), ctx->s_main));
    const DivC d = divc_of(c);
    k_divcheck<2><<<ctx->sm_count * 8, 256, 0, ctx->s_main>>>(d, seed, n / 2, 0, ctx->d_red);
    k_divcheck<2><<<ctx->sm_count * 8, 256, 0, ctx->s_main>>>(d, seed ^ 0xABCDEFull, n - n / 2, 1, ctx->d_red);
    ctx->n_launches += 2;
    CHMY_CUDA(cudaGetLastError());
    CHMY_CUDA(cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_main));
    CHMY_CUDA(cudaStreamSynchronize(ctx->s_main));
    *mismatches = ctx->h_red[0];
    return CHMY_OK;
}

// ---------------------------------------------------------------------------------------------- update_stress! 3D
struct Stress3P {
    double *t[6], *o[6], *Pr, *dV;            // tau, tau_old: xx yy zz xy xz yz ; pointers address logical (0,0,0)
    const double *Vx, *Vy, *Vz;
    Strides cc, vc, cv, vv;                   // CC: xx yy zz Pr dV Vz ; VC: Vx xz ; CV: Vy yz ; VV: xy
    int lo[3], hi[3];                         // box, hi exclusive
    double idx, idy, idz, eta_ve, dtau_Pr, dtau_r;
    DivC Gdt, eta, three;
    int sync, cz;                             // sync: keep the warps of a CTA on the same plane (L1 reuse of neighbour rows)
};

template <bool TD>
__device__ __forceinline__ double stress_upd(double t, double to, double e2, const Stress3P& p) {
    // tau + (((-(tau - tau_old))/(G dt) - tau/eta) + 2 e) * eta_ve * dtau_r      (stokes_3d_inc_ve_T.jl:34-45)
    const double r = (div_u<TD>(-(t - to), p.Gdt) - div_u<TD>(t, p.eta)) + e2;
    return t + (r * p.eta_ve) * p.dtau_r;
}

template <bool TD, int TYB>
__global__ void __launch_bounds__(TX* TYB, 16 / TYB) k_stress3(const Stress3P p) {
    const int lane = threadIdx.x;
    const int i = p.lo[0] + (blockIdx.x * TX + lane) * 2;
    const int j = p.lo[1] + blockIdx.y * TYB + threadIdx.y;
    const int k0 = p.lo[2] + blockIdx.z * p.cz;
    const int k1 = min(k0 + p.cz, p.hi[2]);
    const int nact = j < p.hi[1] ? min(max(p.hi[0] - i, 0), 2) : 0;   // cells of this lane inside the box
    const bool act = nact > 0;
    // the +x neighbour of a lane's second cell lives in the next lane unless that lane is outside the box / warp
    const bool xlast = lane == TX - 1 || i + 2 >= p.hi[0];

    long long cc = (long long)i + (long long)j * p.cc.sy + (long long)k0 * p.cc.sz;
    long long vc = (long long)i + (long long)j * p.vc.sy + (long long)k0 * p.vc.sz;
    long long cv = (long long)i + (long long)j * p.cv.sy + (long long)k0 * p.cv.sz;
    long long vv = (long long)i + (long long)j * p.vv.sy + (long long)k0 * p.vv.sz;

    const double2 z2 = make_double2(0.0, 0.0);
    double2 vx_km = z2, vy_km = z2, vz_k = z2, vzjm = z2;
    if (act) {
        vx_km = ld2(p.Vx + vc - p.vc.sz);
        vy_km = ld2(p.Vy + cv - p.cv.sz);
        vz_k  = ld2(p.Vz + cc);
        vzjm  = ld2(p.Vz + cc - p.cc.sy);
    }
    for (int k = k0; k < k1; ++k) {
        if (p.sync) __syncthreads();
        double2 vx = z2, vxjm = z2, vy = z2, vyjp = z2, vzkp = z2, vzjmkp = z2, pr = z2;
        double2 t[6], o[6];
        double  vx_e = 0.0, vy_e = 0.0, vz_e = 0.0;
        if (act) {
            vx   = ld2(p.Vx + vc);
            vxjm = ld2(p.Vx + vc - p.vc.sy);
            vy   = ld2(p.Vy + cv);
            vyjp = ld2(p.Vy + cv + p.cv.sy);
            vzkp = ld2(p.Vz + cc + p.cc.sz);
            vzjmkp = ld2(p.Vz + cc - p.cc.sy + p.cc.sz);       // same plane as vzkp: the row below loads it in this step
            pr   = ld2(p.Pr + cc);
#pragma unroll
            for (int c = 0; c < 3; ++c) { t[c] = ld2(p.t[c] + cc); o[c] = ld2(p.o[c] + cc); }
            t[3] = ld2(p.t[3] + vv); o[3] = ld2(p.o[3] + vv);
            t[4] = ld2(p.t[4] + vc); o[4] = ld2(p.o[4] + vc);
            t[5] = ld2(p.t[5] + cv); o[5] = ld2(p.o[5] + cv);
            if (xlast) vx_e = p.Vx[vc + 2];                    // Vx[i+2]
            if (lane == 0) { vy_e = p.Vy[cv - 1]; vz_e = p.Vz[cc - 1]; }   // Vy[i-1], Vz[i-1]
        } else {
#pragma unroll
            for (int c = 0; c < 6; ++c) { t[c] = z2; o[c] = z2; }
        }
        // x-neighbours across lanes
        double vx_ip2 = __shfl_down_sync(FULL, vx.x, 1);       // Vx[i+2]
        double vy_im1 = __shfl_up_sync(FULL, vy.y, 1);         // Vy[i-1]
        double vz_im1 = __shfl_up_sync(FULL, vz_k.y, 1);       // Vz[i-1]
        if (xlast) vx_ip2 = vx_e;
        if (lane == 0) { vy_im1 = vy_e; vz_im1 = vz_e; }

        double2 dv, prn, tn[6];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // operands of cell h (0: i, 1: i+1)
            const double a_vx = h ? vx.y : vx.x, a_vxip = h ? vx_ip2 : vx.y, a_vxjm = h ? vxjm.y : vxjm.x, a_vxkm = h ? vx_km.y : vx_km.x;
            const double a_vy = h ? vy.y : vy.x, a_vyjp = h ? vyjp.y : vyjp.x, a_vyim = h ? vy.x : vy_im1, a_vykm = h ? vy_km.y : vy_km.x;
            const double a_vz = h ? vz_k.y : vz_k.x, a_vzkp = h ? vzkp.y : vzkp.x, a_vzim = h ? vz_k.x : vz_im1, a_vzjm = h ? vzjm.y : vzjm.x;
            const double exx = (a_vxip - a_vx) * p.idx;
            const double eyy = (a_vyjp - a_vy) * p.idy;
            const double ezz = (a_vzkp - a_vz) * p.idz;
            const double exy = 0.5 * ((a_vx - a_vxjm) * p.idy + (a_vy - a_vyim) * p.idx);
            const double exz = 0.5 * ((a_vx - a_vxkm) * p.idz + (a_vz - a_vzim) * p.idx);
            const double eyz = 0.5 * ((a_vy - a_vykm) * p.idz + (a_vz - a_vzjm) * p.idy);
            const double d   = (exx + eyy) + ezz;
            const double d3  = div_u<TD>(d, p.three);
            const double a_pr = h ? pr.y : pr.x;
            const double n_pr = a_pr - (d * p.eta_ve) * p.dtau_Pr;
            const double e2[6] = {2.0 * (exx - d3), 2.0 * (eyy - d3), 2.0 * (ezz - d3), 2.0 * exy, 2.0 * exz, 2.0 * eyz};
            if (h) { dv.y = d; prn.y = n_pr; } else { dv.x = d; prn.x = n_pr; }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double r = stress_upd<TD>(h ? t[c].y : t[c].x, h ? o[c].y : o[c].x, e2[c], p);
                if (h) tn[c].y = r; else tn[c].x = r;
            }
        }
        if (nact == 2) {
            st2(p.dV + cc, dv);
            st2(p.Pr + cc, prn);
#pragma unroll
            for (int c = 0; c < 3; ++c) st2(p.t[c] + cc, tn[c]);
            st2(p.t[3] + vv, tn[3]);
            st2(p.t[4] + vc, tn[4]);
            st2(p.t[5] + cv, tn[5]);
        } else if (nact == 1) {
            p.dV[cc] = dv.x;
            p.Pr[cc] = prn.x;
#pragma unroll
            for (int c = 0; c < 3; ++c) p.t[c][cc] = tn[c].x;
            p.t[3][vv] = tn[3].x;
            p.t[4][vc] = tn[4].x;
            p.t[5][cv] = tn[5].x;
        }
        vx_km = vx; vy_km = vy; vz_k = vzkp; vzjm = vzjmkp;
        cc += p.cc.sz; vc += p.vc.sz; cv += p.cv.sz; vv += p.vv.sz;
    }
}

// ---------------------------------------------------------------------------------------------- update_velocity! 3D
struct Velocity3P {
    double *Vx, *Vy, *Vz, *rx, *ry, *rz;
    const double *Pr, *t[6], *rho;            // rho == nullptr -> FunctionField inclusion
    Strides cc, vc, cv, vv;                   // CC: Pr xx yy zz Vz rz rho ; VC: Vx rx xz ; CV: Vy ry yz ; VV: xy
    int lo[3], hi[3];
    double idx, idy, idz, nudtau;
    DivC eta_ve;
    InclDev inc;
    int sync, cz;
};

template <bool TD, bool FUN, int TYB>
__global__ void __launch_bounds__(TX* TYB, 16 / TYB) k_velocity3(const Velocity3P p) {
    const int lane = threadIdx.x;
    const int i = p.lo[0] + (blockIdx.x * TX + lane) * 2;
    const int j = p.lo[1] + blockIdx.y * TYB + threadIdx.y;
    const int k0 = p.lo[2] + blockIdx.z * p.cz;
    const int k1 = min(k0 + p.cz, p.hi[2]);
    const int nact = j < p.hi[1] ? min(max(p.hi[0] - i, 0), 2) : 0;
    const bool act = nact > 0;
    const bool xlast = lane == TX - 1 || i + 2 >= p.hi[0];

    long long cc = (long long)i + (long long)j * p.cc.sy + (long long)k0 * p.cc.sz;
    long long vc = (long long)i + (long long)j * p.vc.sy + (long long)k0 * p.vc.sz;
    long long cv = (long long)i + (long long)j * p.cv.sy + (long long)k0 * p.cv.sz;
    long long vv = (long long)i + (long long)j * p.vv.sy + (long long)k0 * p.vv.sz;

    // FunctionField rho_g at (Center, Center, Vertex): the x and y parts of the radius are per-thread constants
    double sxy0 = 0.0, sxy1 = 0.0;
    if (FUN) {
        const double cy = coord_dev(p.inc.origin[1], p.inc.spacing[1], p.inc.loc[1], j) - p.inc.c0[1];
        const double c0 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], i) - p.inc.c0[0];
        const double c1 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], i + 1) - p.inc.c0[0];
        sxy0 = c0 * c0 + cy * cy;
        sxy1 = c1 * c1 + cy * cy;
    }

    const double2 z2 = make_double2(0.0, 0.0);
    double2 pr_km = z2, tzz_km = z2, txz_k = z2, tyz_k = z2, tyzjp_k = z2;
    if (act) {
        pr_km   = ld2(p.Pr + cc - p.cc.sz);
        tzz_km  = ld2(p.t[2] + cc - p.cc.sz);
        txz_k   = ld2(p.t[4] + vc);
        tyz_k   = ld2(p.t[5] + cv);
        tyzjp_k = ld2(p.t[5] + cv + p.cv.sy);
    }
    for (int k = k0; k < k1; ++k) {
        if (p.sync) __syncthreads();
        double2 pr = z2, prjm = z2, txx = z2, tyy = z2, tyyjm = z2, tzz = z2, txy = z2, txyjp = z2, txzkp = z2, tyzkp = z2,
                tyzjpkp = z2, vx = z2, vy = z2, vz = z2, rho = z2;
        double pr_e = 0.0, txx_e = 0.0, txy_e = 0.0, txz_e = 0.0;
        if (act) {
            pr      = ld2(p.Pr + cc);
            prjm    = ld2(p.Pr + cc - p.cc.sy);
            txx     = ld2(p.t[0] + cc);
            tyy     = ld2(p.t[1] + cc);
            tyyjm   = ld2(p.t[1] + cc - p.cc.sy);
            tzz     = ld2(p.t[2] + cc);
            txy     = ld2(p.t[3] + vv);
            txyjp   = ld2(p.t[3] + vv + p.vv.sy);
            txzkp   = ld2(p.t[4] + vc + p.vc.sz);
            tyzkp   = ld2(p.t[5] + cv + p.cv.sz);
            tyzjpkp = ld2(p.t[5] + cv + p.cv.sy + p.cv.sz);
            vx      = ld2(p.Vx + vc);
            vy      = ld2(p.Vy + cv);
            vz      = ld2(p.Vz + cc);
            if (!FUN) rho = ld2(p.rho + cc);
            if (lane == 0) { pr_e = p.Pr[cc - 1]; txx_e = p.t[0][cc - 1]; }
            if (xlast) { txy_e = p.t[3][vv + 2]; txz_e = p.t[4][vc + 2]; }
        }
        double pr_im1  = __shfl_up_sync(FULL, pr.y, 1);
        double txx_im1 = __shfl_up_sync(FULL, txx.y, 1);
        double txy_ip2 = __shfl_down_sync(FULL, txy.x, 1);
        double txz_ip2 = __shfl_down_sync(FULL, txz_k.x, 1);
        if (lane == 0) { pr_im1 = pr_e; txx_im1 = txx_e; }
        if (xlast) { txy_ip2 = txy_e; txz_ip2 = txz_e; }
        if (FUN) {
            const double cz = coord_dev(p.inc.origin[2], p.inc.spacing[2], p.inc.loc[2], k) - p.inc.c0[2];
            const double cz2 = cz * cz;
            rho.x = (sxy0 + cz2) < p.inc.r2 ? p.inc.in : p.inc.out;
            rho.y = (sxy1 + cz2) < p.inc.r2 ? p.inc.in : p.inc.out;
        }
        double2 nrx, nry, nrz, nvx, nvy, nvz;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a_pr = h ? pr.y : pr.x, a_prim = h ? pr.x : pr_im1, a_prjm = h ? prjm.y : prjm.x, a_prkm = h ? pr_km.y : pr_km.x;
            const double a_txx = h ? txx.y : txx.x, a_txxim = h ? txx.x : txx_im1;
            const double a_tyy = h ? tyy.y : tyy.x, a_tyyjm = h ? tyyjm.y : tyyjm.x;
            const double a_tzz = h ? tzz.y : tzz.x, a_tzzkm = h ? tzz_km.y : tzz_km.x;
            const double a_txy = h ? txy.y : txy.x, a_txyjp = h ? txyjp.y : txyjp.x, a_txyip = h ? txy_ip2 : txy.y;
            const double a_txz = h ? txz_k.y : txz_k.x, a_txzkp = h ? txzkp.y : txzkp.x, a_txzip = h ? txz_ip2 : txz_k.y;
            const double a_tyz = h ? tyz_k.y : tyz_k.x, a_tyzkp = h ? tyzkp.y : tyzkp.x, a_tyzjp = h ? tyzjp_k.y : tyzjp_k.x;
            const double rvx = (((-((a_pr - a_prim) * p.idx)) + (a_txx - a_txxim) * p.idx) + (a_txyjp - a_txy) * p.idy) +
                               (a_txzkp - a_txz) * p.idz;
            const double rvy = (((-((a_pr - a_prjm) * p.idy)) + (a_tyy - a_tyyjm) * p.idy) + (a_txyip - a_txy) * p.idx) +
                               (a_tyzkp - a_tyz) * p.idz;
            const double rvz = ((((-((a_pr - a_prkm) * p.idz)) + (a_tzz - a_tzzkm) * p.idz) + (a_txzip - a_txz) * p.idx) +
                                (a_tyzjp - a_tyz) * p.idy) - (h ? rho.y : rho.x);
            const double ux = (h ? vx.y : vx.x) + div_u<TD>(rvx * p.nudtau, p.eta_ve);
            const double uy = (h ? vy.y : vy.x) + div_u<TD>(rvy * p.nudtau, p.eta_ve);
            const double uz = (h ? vz.y : vz.x) + div_u<TD>(rvz * p.nudtau, p.eta_ve);
            if (h) { nrx.y = rvx; nry.y = rvy; nrz.y = rvz; nvx.y = ux; nvy.y = uy; nvz.y = uz; }
            else   { nrx.x = rvx; nry.x = rvy; nrz.x = rvz; nvx.x = ux; nvy.x = uy; nvz.x = uz; }
        }
        if (nact == 2) {
            st2(p.rx + vc, nrx); st2(p.ry + cv, nry); st2(p.rz + cc, nrz);
            st2(p.Vx + vc, nvx); st2(p.Vy + cv, nvy); st2(p.Vz + cc, nvz);
        } else if (nact == 1) {
            p.rx[vc] = nrx.x; p.ry[cv] = nry.x; p.rz[cc] = nrz.x;
            p.Vx[vc] = nvx.x; p.Vy[cv] = nvy.x; p.Vz[cc] = nvz.x;
        }
        pr_km = pr; tzz_km = tzz; txz_k = txzkp; tyz_k = tyzkp; tyzjp_k = tyzjpkp;
        cc += p.cc.sz; vc += p.vc.sz; cv += p.cv.sz; vv += p.vv.sz;
    }
}

// ---------------------------------------------------------------------------------------------- thermal pair 3D
// update_thermal_flux! (stokes_3d_inc_ve_T.jl:59-71):
//   q.d[I] = ((-lam)*((T[I]-T[I-e_d])*id_d) + max(V.d[I],0)*T[I-e_d]) + min(V.d[I],0)*T[I]
struct Flux3P {
    double *qx, *qy, *qz;
    const double *T, *Vx, *Vy, *Vz;
    Strides cc, vc, cv;                       // CC: T Vz qz ; VC: Vx qx ; CV: Vy qy
    int lo[3], hi[3];
    double lam, idx, idy, idz;
    int sync, cz;
};

__device__ __forceinline__ double flux1(double nlam, double t, double tm, double v, double id) {
    return (nlam * ((t - tm) * id) + jl_max0(v) * tm) + jl_min0(v) * t;
}

__global__ void __launch_bounds__(TX* TY) k_flux3(const Flux3P p) {
    const int lane = threadIdx.x;
    const int i = p.lo[0] + (blockIdx.x * TX + lane) * 2;
    const int j = p.lo[1] + blockIdx.y * TY + threadIdx.y;
    const int k0 = p.lo[2] + blockIdx.z * p.cz;
    const int k1 = min(k0 + p.cz, p.hi[2]);
    const int nact = j < p.hi[1] ? min(max(p.hi[0] - i, 0), 2) : 0;
    const bool act = nact > 0;
    long long cc = (long long)i + (long long)j * p.cc.sy + (long long)k0 * p.cc.sz;
    long long vc = (long long)i + (long long)j * p.vc.sy + (long long)k0 * p.vc.sz;
    long long cv = (long long)i + (long long)j * p.cv.sy + (long long)k0 * p.cv.sz;
    const double2 z2 = make_double2(0.0, 0.0);
    const double nlam = -p.lam;
    double2 T_km = act ? ld2(p.T + cc - p.cc.sz) : z2;
    for (int k = k0; k < k1; ++k) {
        if (p.sync) __syncthreads();
        double2 T = z2, Tjm = z2, vx = z2, vy = z2, vz = z2;
        double e = 0.0;
        if (act) {
            T   = ld2(p.T + cc);
            Tjm = ld2(p.T + cc - p.cc.sy);
            vx  = ld2(p.Vx + vc);
            vy  = ld2(p.Vy + cv);
            vz  = ld2(p.Vz + cc);
            if (lane == 0) e = p.T[cc - 1];
        }
        const double T_im1 = nb_left(T.y, e, lane);
        double2 qx, qy, qz;
        qx.x = flux1(nlam, T.x, T_im1, vx.x, p.idx);
        qx.y = flux1(nlam, T.y, T.x, vx.y, p.idx);
        qy.x = flux1(nlam, T.x, Tjm.x, vy.x, p.idy);
        qy.y = flux1(nlam, T.y, Tjm.y, vy.y, p.idy);
        qz.x = flux1(nlam, T.x, T_km.x, vz.x, p.idz);
        qz.y = flux1(nlam, T.y, T_km.y, vz.y, p.idz);
        if (nact == 2) { st2(p.qx + vc, qx); st2(p.qy + cv, qy); st2(p.qz + cc, qz); }
        else if (nact == 1) { p.qx[vc] = qx.x; p.qy[cv] = qy.x; p.qz[cc] = qz.x; }
        T_km = T;
        cc += p.cc.sz; vc += p.vc.sz; cv += p.cv.sz;
    }
}

// update_thermal! (stokes_3d_inc_ve_T.jl:73-77): T[I] = T_old[I] - dt*(((dx qx) + (dy qy)) + (dz qz))
struct Thermal3P {
    double* T;
    const double *To, *qx, *qy, *qz;
    Strides cc, vc, cv;                       // CC: T To qz ; VC: qx ; CV: qy
    int lo[3], hi[3];
    double dt, idx, idy, idz;
    int sync, cz;
};

__global__ void __launch_bounds__(TX* TY) k_thermal3(const Thermal3P p) {
    const int lane = threadIdx.x;
    const int i = p.lo[0] + (blockIdx.x * TX + lane) * 2;
    const int j = p.lo[1] + blockIdx.y * TY + threadIdx.y;
    const int k0 = p.lo[2] + blockIdx.z * p.cz;
    const int k1 = min(k0 + p.cz, p.hi[2]);
    const int nact = j < p.hi[1] ? min(max(p.hi[0] - i, 0), 2) : 0;
    const bool act = nact > 0;
    const bool xlast = lane == TX - 1 || i + 2 >= p.hi[0];
    long long cc = (long long)i + (long long)j * p.cc.sy + (long long)k0 * p.cc.sz;
    long long vc = (long long)i + (long long)j * p.vc.sy + (long long)k0 * p.vc.sz;
    long long cv = (long long)i + (long long)j * p.cv.sy + (long long)k0 * p.cv.sz;
    const double2 z2 = make_double2(0.0, 0.0);
    double2 qz = act ? ld2(p.qz + cc) : z2;
    for (int k = k0; k < k1; ++k) {
        if (p.sync) __syncthreads();
        double2 to = z2, qx = z2, qy = z2, qyjp = z2, qzkp = z2;
        double e = 0.0;
        if (act) {
            to   = ld2(p.To + cc);
            qx   = ld2(p.qx + vc);
            qy   = ld2(p.qy + cv);
            qyjp = ld2(p.qy + cv + p.cv.sy);
            qzkp = ld2(p.qz + cc + p.cc.sz);
            if (xlast) e = p.qx[vc + 2];
        }
        const double qx_ip2 = nb_right(qx.x, e, xlast);
        double2 o;
        o.x = to.x - p.dt * (((qx.y - qx.x) * p.idx + (qyjp.x - qy.x) * p.idy) + (qzkp.x - qz.x) * p.idz);
        o.y = to.y - p.dt * (((qx_ip2 - qx.y) * p.idx + (qyjp.y - qy.y) * p.idy) + (qzkp.y - qz.y) * p.idz);
        if (nact == 2) st2(p.T + cc, o);
        else if (nact == 1) p.T[cc] = o.x;
        qz = qzkp;
        cc += p.cc.sz; vc += p.vc.sz; cv += p.cv.sz;
    }
}

// ---------------------------------------------------------------------------------------------- dispatch
static bool g_force_true_div = false, g_disable_fast = false, g_env_read = false;
static int  g_ty = 8, g_sync = 1, g_cz = 0;
static void read_env() {
    if (g_env_read) return;
    g_env_read = true;
    const char* a = getenv("CHMY_TRUE_DIV");
    const char* b = getenv("CHMY_NO_FAST");
    const char* h = getenv("CHMY_TY");
    const char* y = getenv("CHMY_SYNC");
    const char* z = getenv("CHMY_CZ");
    if (h) g_ty = atoi(h) == 16 ? 16 : 8;
    if (y) g_sync = atoi(y);
    if (z) g_cz = atoi(z);
    g_force_true_div = a && a[0] == '1';
    g_disable_fast   = b && b[0] == '1';
}

extern "C" int chmy_set_tuning(int disable_fast_kernels, int force_true_division) {
    read_env();
    if (disable_fast_kernels >= 0) g_disable_fast = disable_fast_kernels != 0;
    if (force_true_division >= 0) g_force_true_div = force_true_division != 0;
    return CHMY_OK;
}

bool chmy_fast_disabled() { read_env(); return g_disable_fast; }
bool chmy_force_true_div() { read_env(); return g_force_true_div; }

// z-chunk length: balanced chunks of about `target` planes
static int pick_cz(int nz, int target) {
    const int nch = (nz + target - 1) / target;
    return (nz + nch - 1) / nch;
}
static dim3 march_grid2(const Box& b, int ty, int cz) {
    return dim3((unsigned)((b.n[0] + 2 * TX - 1) / (2 * TX)), (unsigned)((b.n[1] + ty - 1) / ty), (unsigned)((b.n[2] + cz - 1) / cz));
}

int chmy_run_op_fast(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st, int* handled) {
    *handled = 0;
    read_env();
    if (g_disable_fast) return CHMY_OK;
    const int nd = d->grid.ndims;
    chmy_field* const* F = d->fields;
    const double* s = d->scalars;
    const double* id = d->grid.inv_spacing;
    if (nd == 2) return chmy_run_op_fast2d(ctx, d, box, st, handled);
    if (nd != 3 || (box.lo[0] & 1)) return CHMY_OK;
    for (int q = 0; q < d->nfields; ++q)
        if (F[q] && !aligned16(F[q])) return CHMY_OK;

    if (d->op == CHMY_OP_UPDATE_THERMAL_FLUX) {     // fields: qT.x qT.y qT.z T V.x V.y V.z ; scalars: lambda
        if (!same_strides(F[0], F[4]) || !same_strides(F[1], F[5]) || !same_strides(F[2], F[3]) || !same_strides(F[6], F[3]))
            return CHMY_OK;
        Flux3P p;
        p.qx = F[0]->p0; p.qy = F[1]->p0; p.qz = F[2]->p0; p.T = F[3]->p0;
        p.Vx = F[4]->p0; p.Vy = F[5]->p0; p.Vz = F[6]->p0;
        p.cc = strides_of(F[3]); p.vc = strides_of(F[0]); p.cv = strides_of(F[1]);
        for (int a = 0; a < 3; ++a) { p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a]; }
        p.lam = s[0]; p.idx = id[0]; p.idy = id[1]; p.idz = id[2];
        p.sync = g_sync; p.cz = pick_cz(box.n[2], g_cz > 0 ? g_cz : CZ);
        k_flux3<<<march_grid2(box, TY, p.cz), dim3(TX, TY, 1), 0, st>>>(p);
        ctx->n_launches++;
        CHMY_CUDA(cudaGetLastError());
        *handled = 1;
        return CHMY_OK;
    }
    if (d->op == CHMY_OP_UPDATE_THERMAL) {          // fields: T T_old qT.x qT.y qT.z ; scalars: dt
        if (!same_strides(F[0], F[1]) || !same_strides(F[4], F[0])) return CHMY_OK;
        Thermal3P p;
        p.T = F[0]->p0; p.To = F[1]->p0; p.qx = F[2]->p0; p.qy = F[3]->p0; p.qz = F[4]->p0;
        p.cc = strides_of(F[0]); p.vc = strides_of(F[2]); p.cv = strides_of(F[3]);
        for (int a = 0; a < 3; ++a) { p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a]; }
        p.dt = s[0]; p.idx = id[0]; p.idy = id[1]; p.idz = id[2];
        p.sync = g_sync; p.cz = pick_cz(box.n[2], g_cz > 0 ? g_cz : CZ);
        k_thermal3<<<march_grid2(box, TY, p.cz), dim3(TX, TY, 1), 0, st>>>(p);
        ctx->n_launches++;
        CHMY_CUDA(cudaGetLastError());
        *handled = 1;
        return CHMY_OK;
    }

    if (d->op == CHMY_OP_UPDATE_STRESS) {
        // storage classes must be consistent (they are, for fields created on the same grid)
        const chmy_field *CC = F[0], *VV = F[3], *VC = F[4], *CV = F[5];
        if (!same_strides(F[1], CC) || !same_strides(F[2], CC) || !same_strides(F[6], CC) || !same_strides(F[7], CC) ||
            !same_strides(F[10], CC) || !same_strides(F[8], VC) || !same_strides(F[9], CV))
            return CHMY_OK;
        for (int c = 0; c < 6; ++c)
            if (!same_strides(F[11 + c], F[c])) return CHMY_OK;
        Stress3P p;
        for (int c = 0; c < 6; ++c) { p.t[c] = F[c]->p0; p.o[c] = F[11 + c]->p0; }
        p.Pr = F[6]->p0; p.dV = F[7]->p0;
        p.Vx = F[8]->p0; p.Vy = F[9]->p0; p.Vz = F[10]->p0;
        p.cc = strides_of(CC); p.vc = strides_of(VC); p.cv = strides_of(CV); p.vv = strides_of(VV);
        for (int a = 0; a < 3; ++a) { p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a]; }
        p.idx = id[0]; p.idy = id[1]; p.idz = id[2];
        p.eta_ve = s[1]; p.dtau_Pr = s[4]; p.dtau_r = s[5];
        const double Gdt = s[2] * s[3];
        p.Gdt = divc_of(Gdt); p.eta = divc_of(s[0]); p.three = divc_of(3.0);
        const bool td = g_force_true_div || !markstein_ok(Gdt) || !markstein_ok(s[0]);
        const dim3 blk(TX, TY, 1);
        p.sync = g_sync; p.cz = pick_cz(box.n[2], g_cz > 0 ? g_cz : CZ);
        if (g_ty == 16) {
            const dim3 b16(TX, 16, 1), g16 = march_grid2(box, 16, p.cz);
            if (td) k_stress3<true, 16><<<g16, b16, 0, st>>>(p);
            else k_stress3<false, 16><<<g16, b16, 0, st>>>(p);
        } else {
            const dim3 g8 = march_grid2(box, 8, p.cz);
            if (td) k_stress3<true, 8><<<g8, blk, 0, st>>>(p);
            else k_stress3<false, 8><<<g8, blk, 0, st>>>(p);
        }
        ctx->n_launches++;
        CHMY_CUDA(cudaGetLastError());
        *handled = 1;
        return CHMY_OK;
    }
    if (d->op == CHMY_OP_UPDATE_VELOCITY) {
        const chmy_field *CC = F[6], *VV = F[10], *VC = F[11], *CV = F[12], *rho = F[13];
        if (!same_strides(F[7], CC) || !same_strides(F[8], CC) || !same_strides(F[9], CC) || !same_strides(F[2], CC) ||
            !same_strides(F[5], CC) || !same_strides(F[0], VC) || !same_strides(F[3], VC) || !same_strides(F[1], CV) ||
            !same_strides(F[4], CV) || (rho && !same_strides(rho, CC)))
            return CHMY_OK;
        Velocity3P p;
        p.Vx = F[0]->p0; p.Vy = F[1]->p0; p.Vz = F[2]->p0;
        p.rx = F[3]->p0; p.ry = F[4]->p0; p.rz = F[5]->p0;
        p.Pr = F[6]->p0;
        for (int c = 0; c < 6; ++c) p.t[c] = F[7 + c]->p0;
        p.rho = rho ? rho->p0 : nullptr;
        p.cc = strides_of(CC); p.vc = strides_of(VC); p.cv = strides_of(CV); p.vv = strides_of(VV);
        for (int a = 0; a < 3; ++a) { p.lo[a] = box.lo[a]; p.hi[a] = box.lo[a] + box.n[a]; }
        p.idx = id[0]; p.idy = id[1]; p.idz = id[2];
        p.nudtau = s[1];
        p.eta_ve = divc_of(s[0]);
        memset(&p.inc, 0, sizeof(p.inc));
        if (!rho) {
            p.inc.active = 1; p.inc.nd = 3;
            for (int a = 0; a < 3; ++a) {
                p.inc.loc[a] = d->rho_g.loc[a]; p.inc.origin[a] = d->grid.origin[a];
                p.inc.spacing[a] = d->grid.spacing[a]; p.inc.c0[a] = d->rho_g.c0[a];
            }
            p.inc.r2 = d->rho_g.r * d->rho_g.r; p.inc.in = d->rho_g.in; p.inc.out = d->rho_g.out;
        }
        const bool td = g_force_true_div || !markstein_ok(s[0]);
        const dim3 blk(TX, TY, 1);
        p.sync = g_sync; p.cz = pick_cz(box.n[2], g_cz > 0 ? g_cz : CZ);
        if (g_ty == 16) {
            const dim3 b16(TX, 16, 1), g16 = march_grid2(box, 16, p.cz);
            if (rho) {
                if (td) k_velocity3<true, false, 16><<<g16, b16, 0, st>>>(p);
                else k_velocity3<false, false, 16><<<g16, b16, 0, st>>>(p);
            } else {
                if (td) k_velocity3<true, true, 16><<<g16, b16, 0, st>>>(p);
                else k_velocity3<false, true, 16><<<g16, b16, 0, st>>>(p);
            }
        } else {
            const dim3 g8 = march_grid2(box, 8, p.cz);
            if (rho) {
                if (td) k_velocity3<true, false, 8><<<g8, blk, 0, st>>>(p);
                else k_velocity3<false, false, 8><<<g8, blk, 0, st>>>(p);
            } else {
                if (td) k_velocity3<true, true, 8><<<g8, blk, 0, st>>>(p);
                else k_velocity3<false, true, 8><<<g8, blk, 0, st>>>(p);
            }
        }
        ctx->n_launches++;
        CHMY_CUDA(cudaGetLastError());
        *handled = 1;
        return CHMY_OK;
    }
    return CHMY_OK;
}
