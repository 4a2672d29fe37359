// ops_fast2d.cu -- tuned 2D kernels: diffusion (examples/diffusion_2d.jl:8-19), 2D Stokes
// (examples/stokes_2d_inc_ve_T.jl:20-43) and the 2D thermal pair (:45-60).
//
// Same recipe as the 3D kernels of ops_fast.cu, one dimension down: pitched fields (logical x-index 0 of every row on
// a 128-byte boundary), a lane owns an aligned pair of cells (i, i+1) and moves every array with 128-bit accesses,
// x-neighbours come from warp shuffles, and the thread marches along y keeping the j-1 / j+1 rows of the stencil
// operands in registers, so every array element is requested from DRAM once.  A CTA is 4 warps side by side in x
// (256 cells) marching CY rows.
//
// Arithmetic order is exactly that of ops.cu / the reference (compiled with -fmad=false; fma() only where written).
#include "fast_common.cuh"

constexpr int WX = 4;      // warps per CTA, side by side in x
constexpr int CY = 16;     // target rows marched by one CTA (measured optimum: short chunks, many CTAs)

struct Geo2 {
    int lo[2], hi[2];      // box, hi exclusive
    int cy;                // rows marched by one CTA
};

// per-thread geometry; returns false when the whole warp lies outside the box (warp-uniform)
struct Lane2 {
    int  lane, i, j0, j1, nact;
    bool act, xlast;
    __device__ __forceinline__ bool init(const Geo2& g) {
        lane = threadIdx.x;
        const int seg = blockIdx.x * WX + threadIdx.y;
        const int i0  = g.lo[0] + seg * 64;
        if (i0 >= g.hi[0]) return false;
        i     = i0 + lane * 2;
        j0    = g.lo[1] + blockIdx.y * g.cy;
        j1    = min(j0 + g.cy, g.hi[1]);
        nact  = min(max(g.hi[0] - i, 0), 2);
        act   = nact > 0;
        xlast = lane == TX - 1 || i + 2 >= g.hi[0];
        return true;
    }
};

__device__ __forceinline__ void store_pair(double* p, double2 v, int nact) {
    if (nact == 2) st2(p, v);
    else if (nact == 1) p[0] = v.x;
}

static dim3 grid2(const Box& b, int cy) {
    return dim3((unsigned)((b.n[0] + 64 * WX - 1) / (64 * WX)), (unsigned)((b.n[1] + cy - 1) / cy), 1);
}

// ---------------------------------------------------------------------------------------------- compute_q!
// qx[i,j] = (-chi)*((C[i,j]-C[i-1,j])*idx) ; qy[i,j] = (-chi)*((C[i,j]-C[i,j-1])*idy)
struct DiffQP {
    double *qx, *qy;
    const double* C;
    int    s_cc, s_vc, s_cv;   // row strides of C, qx, qy
    Geo2   g;
    double chi, idx, idy;
};

__global__ void __launch_bounds__(TX* WX) k_diff_q(const DiffQP p) {
    Lane2 t;
    if (!t.init(p.g)) return;
    const double2 z2 = make_double2(0.0, 0.0);
    long long cc = (long long)t.i + (long long)t.j0 * p.s_cc;
    long long vc = (long long)t.i + (long long)t.j0 * p.s_vc;
    long long cv = (long long)t.i + (long long)t.j0 * p.s_cv;
    double2 c_jm = t.act ? ld2(p.C + cc - p.s_cc) : z2;
    const double nchi = -p.chi;
#pragma unroll 2
    for (int j = t.j0; j < t.j1; ++j) {
        double2 c = z2;
        double  e = 0.0;
        if (t.act) {
            c = ld2(p.C + cc);
            if (t.lane == 0) e = p.C[cc - 1];
        }
        const double c_im1 = nb_left(c.y, e, t.lane);
        double2 qx, qy;
        qx.x = nchi * ((c.x - c_im1) * p.idx);
        qx.y = nchi * ((c.y - c.x) * p.idx);
        qy.x = nchi * ((c.x - c_jm.x) * p.idy);
        qy.y = nchi * ((c.y - c_jm.y) * p.idy);
        store_pair(p.qx + vc, qx, t.nact);
        store_pair(p.qy + cv, qy, t.nact);
        c_jm = c;
        cc += p.s_cc; vc += p.s_vc; cv += p.s_cv;
    }
}

// ---------------------------------------------------------------------------------------------- update_C! / update_thermal! (2D)
// dst[i,j] = src[i,j] - dt*((qx[i+1,j]-qx[i,j])*idx + (qy[i,j+1]-qy[i,j])*idy)      (src == dst for update_C!)
struct DivQ2P {
    double*       dst;
    const double *src, *qx, *qy;
    int    s_cc, s_vc, s_cv;
    Geo2   g;
    double dt, idx, idy;
};

__global__ void __launch_bounds__(TX* WX) k_divq2(const DivQ2P p) {
    Lane2 t;
    if (!t.init(p.g)) return;
    const double2 z2 = make_double2(0.0, 0.0);
    long long cc = (long long)t.i + (long long)t.j0 * p.s_cc;
    long long vc = (long long)t.i + (long long)t.j0 * p.s_vc;
    long long cv = (long long)t.i + (long long)t.j0 * p.s_cv;
    double2 qy = t.act ? ld2(p.qy + cv) : z2;
#pragma unroll 2
    for (int j = t.j0; j < t.j1; ++j) {
        double2 s = z2, qx = z2, qyjp = z2;
        double  e = 0.0;
        if (t.act) {
            s    = ld2(p.src + cc);
            qx   = ld2(p.qx + vc);
            qyjp = ld2(p.qy + cv + p.s_cv);
            if (t.xlast) e = p.qx[vc + 2];
        }
        const double qx_ip2 = nb_right(qx.x, e, t.xlast);
        double2 o;
        o.x = s.x - p.dt * ((qx.y - qx.x) * p.idx + (qyjp.x - qy.x) * p.idy);
        o.y = s.y - p.dt * ((qx_ip2 - qx.y) * p.idx + (qyjp.y - qy.y) * p.idy);
        store_pair(p.dst + cc, o, t.nact);
        qy = qyjp;
        cc += p.s_cc; vc += p.s_vc; cv += p.s_cv;
    }
}

// ---------------------------------------------------------------------------------------------- update_thermal_flux! (2D)
// q.d[I] = ((-lam)*((T[I]-T[I-e_d])*id_d) + max(V.d[I],0)*T[I-e_d]) + min(V.d[I],0)*T[I]
struct Flux2P {
    double *qx, *qy;
    const double *T, *Vx, *Vy;
    int    s_cc, s_vc, s_cv;
    Geo2   g;
    double lam, idx, idy;
};

__device__ __forceinline__ double flux1(double nlam, double t, double tm, double v, double id) {
    return (nlam * ((t - tm) * id) + jl_max0(v) * tm) + jl_min0(v) * t;
}

__global__ void __launch_bounds__(TX* WX) k_flux2(const Flux2P p) {
    Lane2 t;
    if (!t.init(p.g)) return;
    const double2 z2 = make_double2(0.0, 0.0);
    long long cc = (long long)t.i + (long long)t.j0 * p.s_cc;
    long long vc = (long long)t.i + (long long)t.j0 * p.s_vc;
    long long cv = (long long)t.i + (long long)t.j0 * p.s_cv;
    double2 T_jm = t.act ? ld2(p.T + cc - p.s_cc) : z2;
    const double nlam = -p.lam;
#pragma unroll 2
    for (int j = t.j0; j < t.j1; ++j) {
        double2 T = z2, vx = z2, vy = z2;
        double  e = 0.0;
        if (t.act) {
            T  = ld2(p.T + cc);
            vx = ld2(p.Vx + vc);
            vy = ld2(p.Vy + cv);
            if (t.lane == 0) e = p.T[cc - 1];
        }
        const double T_im1 = nb_left(T.y, e, t.lane);
        double2 qx, qy;
        qx.x = flux1(nlam, T.x, T_im1, vx.x, p.idx);
        qx.y = flux1(nlam, T.y, T.x, vx.y, p.idx);
        qy.x = flux1(nlam, T.x, T_jm.x, vy.x, p.idy);
        qy.y = flux1(nlam, T.y, T_jm.y, vy.y, p.idy);
        store_pair(p.qx + vc, qx, t.nact);
        store_pair(p.qy + cv, qy, t.nact);
        T_jm = T;
        cc += p.s_cc; vc += p.s_vc; cv += p.s_cv;
    }
}

// ---------------------------------------------------------------------------------------------- update_stress! (2D)
struct Stress2P {
    double *txx, *tyy, *txy, *Pr, *dV;
    const double *Vx, *Vy, *oxx, *oyy, *oxy;
    int    s_cc, s_vc, s_cv, s_vv;     // CC: txx tyy Pr dV oxx oyy ; VC: Vx ; CV: Vy ; VV: txy oxy
    Geo2   g;
    double idx, idy, eta_ve, dtau_Pr, dtau_r;
    DivC   Gdt, eta, three;
};

template <bool TD>
__device__ __forceinline__ double stress_upd2(double t, double to, double e2, const Stress2P& p) {
    const double r = (div_u<TD>(-(t - to), p.Gdt) - div_u<TD>(t, p.eta)) + e2;
    return t + (r * p.eta_ve) * p.dtau_r;
}

template <bool TD>
__global__ void __launch_bounds__(TX* WX) k_stress2(const Stress2P p) {
    Lane2 t;
    if (!t.init(p.g)) return;
    const double2 z2 = make_double2(0.0, 0.0);
    long long cc = (long long)t.i + (long long)t.j0 * p.s_cc;
    long long vc = (long long)t.i + (long long)t.j0 * p.s_vc;
    long long cv = (long long)t.i + (long long)t.j0 * p.s_cv;
    long long vv = (long long)t.i + (long long)t.j0 * p.s_vv;
    double2 vx_jm = z2, vy = z2;
    if (t.act) {
        vx_jm = ld2(p.Vx + vc - p.s_vc);
        vy    = ld2(p.Vy + cv);
    }
    for (int j = t.j0; j < t.j1; ++j) {
        double2 vx = z2, vyjp = z2, pr = z2, a = z2, b = z2, c = z2, oa = z2, ob = z2, oc = z2;
        double  vx_e = 0.0, vy_e = 0.0;
        if (t.act) {
            vx   = ld2(p.Vx + vc);
            vyjp = ld2(p.Vy + cv + p.s_cv);
            pr   = ld2(p.Pr + cc);
            a = ld2(p.txx + cc); oa = ld2(p.oxx + cc);
            b = ld2(p.tyy + cc); ob = ld2(p.oyy + cc);
            c = ld2(p.txy + vv); oc = ld2(p.oxy + vv);
            if (t.xlast) vx_e = p.Vx[vc + 2];
            if (t.lane == 0) vy_e = p.Vy[cv - 1];
        }
        const double vx_ip2 = nb_right(vx.x, vx_e, t.xlast);
        const double vy_im1 = nb_left(vy.y, vy_e, t.lane);
        double2 dv, prn, na, nb, nc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a_vx = h ? vx.y : vx.x, a_vxip = h ? vx_ip2 : vx.y, a_vxjm = h ? vx_jm.y : vx_jm.x;
            const double a_vy = h ? vy.y : vy.x, a_vyjp = h ? vyjp.y : vyjp.x, a_vyim = h ? vy.x : vy_im1;
            const double exx = (a_vxip - a_vx) * p.idx;
            const double eyy = (a_vyjp - a_vy) * p.idy;
            const double exy = 0.5 * ((a_vx - a_vxjm) * p.idy + (a_vy - a_vyim) * p.idx);
            const double d   = exx + eyy;
            const double n_pr = (h ? pr.y : pr.x) - (d * p.eta_ve) * p.dtau_Pr;
            const double d3  = div_u<TD>(d, p.three);          // the 2D driver also divides by 3.0 (:28-29)
            const double ra = stress_upd2<TD>(h ? a.y : a.x, h ? oa.y : oa.x, 2.0 * (exx - d3), p);
            const double rb = stress_upd2<TD>(h ? b.y : b.x, h ? ob.y : ob.x, 2.0 * (eyy - d3), p);
            const double rc = stress_upd2<TD>(h ? c.y : c.x, h ? oc.y : oc.x, 2.0 * exy, p);
            if (h) { dv.y = d; prn.y = n_pr; na.y = ra; nb.y = rb; nc.y = rc; }
            else   { dv.x = d; prn.x = n_pr; na.x = ra; nb.x = rb; nc.x = rc; }
        }
        store_pair(p.dV + cc, dv, t.nact);
        store_pair(p.Pr + cc, prn, t.nact);
        store_pair(p.txx + cc, na, t.nact);
        store_pair(p.tyy + cc, nb, t.nact);
        store_pair(p.txy + vv, nc, t.nact);
        vx_jm = vx; vy = vyjp;
        cc += p.s_cc; vc += p.s_vc; cv += p.s_cv; vv += p.s_vv;
    }
}

// ---------------------------------------------------------------------------------------------- update_velocity! (2D)
struct Velocity2P {
    double *Vx, *Vy, *rx, *ry;
    const double *Pr, *txx, *tyy, *txy, *rho;     // rho == nullptr -> FunctionField inclusion at (Center, Vertex)
    int    s_cc, s_vc, s_cv, s_vv;                // CC: Pr txx tyy ; VC: Vx rx ; CV: Vy ry rho ; VV: txy
    Geo2   g;
    double idx, idy, nudtau;
    DivC   eta_ve;
    InclDev inc;
};

template <bool TD, bool FUN>
__global__ void __launch_bounds__(TX* WX) k_velocity2(const Velocity2P p) {
    Lane2 t;
    if (!t.init(p.g)) return;
    const double2 z2 = make_double2(0.0, 0.0);
    long long cc = (long long)t.i + (long long)t.j0 * p.s_cc;
    long long vc = (long long)t.i + (long long)t.j0 * p.s_vc;
    long long cv = (long long)t.i + (long long)t.j0 * p.s_cv;
    long long vv = (long long)t.i + (long long)t.j0 * p.s_vv;
    double sx0 = 0.0, sx1 = 0.0;
    if (FUN) {
        const double c0 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], t.i) - p.inc.c0[0];
        const double c1 = coord_dev(p.inc.origin[0], p.inc.spacing[0], p.inc.loc[0], t.i + 1) - p.inc.c0[0];
        sx0 = c0 * c0;
        sx1 = c1 * c1;
    }
    double2 pr_jm = z2, tyy_jm = z2, txy = z2;
    if (t.act) {
        pr_jm  = ld2(p.Pr + cc - p.s_cc);
        tyy_jm = ld2(p.tyy + cc - p.s_cc);
        txy    = ld2(p.txy + vv);
    }
    for (int j = t.j0; j < t.j1; ++j) {
        double2 pr = z2, txx = z2, tyy = z2, txyjp = z2, vx = z2, vy = z2, rho = z2;
        double  pr_e = 0.0, txx_e = 0.0, txy_e = 0.0;
        if (t.act) {
            pr    = ld2(p.Pr + cc);
            txx   = ld2(p.txx + cc);
            tyy   = ld2(p.tyy + cc);
            txyjp = ld2(p.txy + vv + p.s_vv);
            vx    = ld2(p.Vx + vc);
            vy    = ld2(p.Vy + cv);
            if (!FUN) rho = ld2(p.rho + cv);
            if (t.lane == 0) { pr_e = p.Pr[cc - 1]; txx_e = p.txx[cc - 1]; }
            if (t.xlast) txy_e = p.txy[vv + 2];
        }
        const double pr_im1  = nb_left(pr.y, pr_e, t.lane);
        const double txx_im1 = nb_left(txx.y, txx_e, t.lane);
        const double txy_ip2 = nb_right(txy.x, txy_e, t.xlast);
        if (FUN) {
            const double cy  = coord_dev(p.inc.origin[1], p.inc.spacing[1], p.inc.loc[1], j) - p.inc.c0[1];
            const double cy2 = cy * cy;
            rho.x = (sx0 + cy2) < p.inc.r2 ? p.inc.in : p.inc.out;
            rho.y = (sx1 + cy2) < p.inc.r2 ? p.inc.in : p.inc.out;
        }
        double2 nrx, nry, nvx, nvy;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a_pr = h ? pr.y : pr.x, a_prim = h ? pr.x : pr_im1, a_prjm = h ? pr_jm.y : pr_jm.x;
            const double a_txx = h ? txx.y : txx.x, a_txxim = h ? txx.x : txx_im1;
            const double a_tyy = h ? tyy.y : tyy.x, a_tyyjm = h ? tyy_jm.y : tyy_jm.x;
            const double a_txy = h ? txy.y : txy.x, a_txyjp = h ? txyjp.y : txyjp.x, a_txyip = h ? txy_ip2 : txy.y;
            const double rvx = ((-((a_pr - a_prim) * p.idx)) + (a_txx - a_txxim) * p.idx) + (a_txyjp - a_txy) * p.idy;
            const double rvy = (((-((a_pr - a_prjm) * p.idy)) + (a_tyy - a_tyyjm) * p.idy) + (a_txyip - a_txy) * p.idx) -
                               (h ? rho.y : rho.x);
            const double ux = (h ? vx.y : vx.x) + div_u<TD>(rvx * p.nudtau, p.eta_ve);
            const double uy = (h ? vy.y : vy.x) + div_u<TD>(rvy * p.nudtau, p.eta_ve);
            if (h) { nrx.y = rvx; nry.y = rvy; nvx.y = ux; nvy.y = uy; }
            else   { nrx.x = rvx; nry.x = rvy; nvx.x = ux; nvy.x = uy; }
        }
        store_pair(p.rx + vc, nrx, t.nact);
        store_pair(p.ry + cv, nry, t.nact);
        store_pair(p.Vx + vc, nvx, t.nact);
        store_pair(p.Vy + cv, nvy, t.nact);
        pr_jm = pr; tyy_jm = tyy; txy = txyjp;
        cc += p.s_cc; vc += p.s_vc; cv += p.s_cv; vv += p.s_vv;
    }
}

// ---------------------------------------------------------------------------------------------- dispatch
static Geo2 geo_of(const Box& b) {
    Geo2 g;
    for (int a = 0; a < 2; ++a) { g.lo[a] = b.lo[a]; g.hi[a] = b.lo[a] + b.n[a]; }
    static int target = 0;
    if (!target) { const char* e = getenv("CHMY_CY"); target = e && atoi(e) > 0 ? atoi(e) : CY; }
    const int nch = (b.n[1] + target - 1) / target;      // balanced chunks of about `target` rows
    g.cy = (b.n[1] + nch - 1) / nch;
    return g;
}

static bool same_sy(const chmy_field* a, const chmy_field* b) { return a->stride[1] == b->stride[1]; }

#define LAUNCH2(P, ...)                                               \
    do {                                                              \
        __VA_ARGS__<<<grid2(box, g.cy), dim3(TX, WX, 1), 0, st>>>(P);       \
        ctx->n_launches++;                                            \
        CHMY_CUDA(cudaGetLastError());                                \
        *handled = 1;                                                 \
        return CHMY_OK;                                               \
    } while (0)

int chmy_run_op_fast2d(chmy_ctx* ctx, const chmy_launch_desc* d, const Box& box, cudaStream_t st, int* handled) {
    *handled = 0;
    chmy_field* const* F = d->fields;
    const double* s  = d->scalars;
    const double* id = d->grid.inv_spacing;
    if (d->grid.ndims != 2 || (box.lo[0] & 1)) return CHMY_OK;
    for (int q = 0; q < d->nfields; ++q)
        if (F[q] && !aligned16(F[q])) return CHMY_OK;
    const Geo2 g = geo_of(box);

    switch (d->op) {
    case CHMY_OP_COMPUTE_Q: {      // fields: q.x q.y C ; scalars: chi
        DiffQP p;
        p.qx = F[0]->p0; p.qy = F[1]->p0; p.C = F[2]->p0;
        p.s_vc = (int)F[0]->stride[1]; p.s_cv = (int)F[1]->stride[1]; p.s_cc = (int)F[2]->stride[1];
        p.g = g; p.chi = s[0]; p.idx = id[0]; p.idy = id[1];
        LAUNCH2(p, k_diff_q);
    }
    case CHMY_OP_UPDATE_C: {       // fields: C q.x q.y ; scalars: dt
        DivQ2P p;
        p.dst = F[0]->p0; p.src = F[0]->p0; p.qx = F[1]->p0; p.qy = F[2]->p0;
        p.s_cc = (int)F[0]->stride[1]; p.s_vc = (int)F[1]->stride[1]; p.s_cv = (int)F[2]->stride[1];
        p.g = g; p.dt = s[0]; p.idx = id[0]; p.idy = id[1];
        LAUNCH2(p, k_divq2);
    }
    case CHMY_OP_UPDATE_THERMAL: { // fields: T T_old qT.x qT.y ; scalars: dt
        if (!same_sy(F[0], F[1])) return CHMY_OK;
        DivQ2P p;
        p.dst = F[0]->p0; p.src = F[1]->p0; p.qx = F[2]->p0; p.qy = F[3]->p0;
        p.s_cc = (int)F[0]->stride[1]; p.s_vc = (int)F[2]->stride[1]; p.s_cv = (int)F[3]->stride[1];
        p.g = g; p.dt = s[0]; p.idx = id[0]; p.idy = id[1];
        LAUNCH2(p, k_divq2);
    }
    case CHMY_OP_UPDATE_THERMAL_FLUX: {   // fields: qT.x qT.y T V.x V.y ; scalars: lambda
        if (!same_sy(F[0], F[3]) || !same_sy(F[1], F[4])) return CHMY_OK;
        Flux2P p;
        p.qx = F[0]->p0; p.qy = F[1]->p0; p.T = F[2]->p0; p.Vx = F[3]->p0; p.Vy = F[4]->p0;
        p.s_vc = (int)F[0]->stride[1]; p.s_cv = (int)F[1]->stride[1]; p.s_cc = (int)F[2]->stride[1];
        p.g = g; p.lam = s[0]; p.idx = id[0]; p.idy = id[1];
        LAUNCH2(p, k_flux2);
    }
    case CHMY_OP_UPDATE_STRESS: {  // fields: txx tyy txy Pr dV Vx Vy oxx oyy oxy ; scalars: eta eta_ve G dt dtau_Pr dtau_r
        const chmy_field* CC = F[0];
        if (!same_sy(F[1], CC) || !same_sy(F[3], CC) || !same_sy(F[4], CC) || !same_sy(F[7], CC) || !same_sy(F[8], CC) ||
            !same_sy(F[9], F[2]))
            return CHMY_OK;
        Stress2P p;
        p.txx = F[0]->p0; p.tyy = F[1]->p0; p.txy = F[2]->p0; p.Pr = F[3]->p0; p.dV = F[4]->p0;
        p.Vx = F[5]->p0; p.Vy = F[6]->p0; p.oxx = F[7]->p0; p.oyy = F[8]->p0; p.oxy = F[9]->p0;
        p.s_cc = (int)CC->stride[1]; p.s_vv = (int)F[2]->stride[1]; p.s_vc = (int)F[5]->stride[1]; p.s_cv = (int)F[6]->stride[1];
        p.g = g; p.idx = id[0]; p.idy = id[1];
        p.eta_ve = s[1]; p.dtau_Pr = s[4]; p.dtau_r = s[5];
        const double Gdt = s[2] * s[3];
        p.Gdt = divc_of(Gdt); p.eta = divc_of(s[0]); p.three = divc_of(3.0);
        const bool td = chmy_force_true_div() || !markstein_ok(Gdt) || !markstein_ok(s[0]);
        if (td) LAUNCH2(p, k_stress2<true>);
        LAUNCH2(p, k_stress2<false>);
    }
    case CHMY_OP_UPDATE_VELOCITY: {  // fields: Vx Vy rx ry Pr txx tyy txy rho|NULL ; scalars: eta_ve nudtau
        const chmy_field *CC = F[4], *rho = F[8];
        if (!same_sy(F[5], CC) || !same_sy(F[6], CC) || !same_sy(F[2], F[0]) || !same_sy(F[3], F[1]) ||
            (rho && !same_sy(rho, F[1])))
            return CHMY_OK;
        Velocity2P p;
        p.Vx = F[0]->p0; p.Vy = F[1]->p0; p.rx = F[2]->p0; p.ry = F[3]->p0;
        p.Pr = F[4]->p0; p.txx = F[5]->p0; p.tyy = F[6]->p0; p.txy = F[7]->p0;
        p.rho = rho ? rho->p0 : nullptr;
        p.s_cc = (int)CC->stride[1]; p.s_vc = (int)F[0]->stride[1]; p.s_cv = (int)F[1]->stride[1]; p.s_vv = (int)F[7]->stride[1];
        p.g = g; p.idx = id[0]; p.idy = id[1]; p.nudtau = s[1];
        p.eta_ve = divc_of(s[0]);
        memset(&p.inc, 0, sizeof(p.inc));
        if (!rho) {
            p.inc.active = 1; p.inc.nd = 2;
            for (int a = 0; a < 2; ++a) {
                p.inc.loc[a] = d->rho_g.loc[a]; p.inc.origin[a] = d->grid.origin[a];
                p.inc.spacing[a] = d->grid.spacing[a]; p.inc.c0[a] = d->rho_g.c0[a];
            }
            p.inc.r2 = d->rho_g.r * d->rho_g.r; p.inc.in = d->rho_g.in; p.inc.out = d->rho_g.out;
        }
        const bool td = chmy_force_true_div() || !markstein_ok(s[0]);
        if (rho) {
            if (td) LAUNCH2(p, k_velocity2<true, false>);
            LAUNCH2(p, k_velocity2<false, false>);
        }
        if (td) LAUNCH2(p, k_velocity2<true, true>);
        LAUNCH2(p, k_velocity2<false, true>);
    }
    default: return CHMY_OK;
    }
}
