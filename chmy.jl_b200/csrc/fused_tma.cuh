// fused_tma.cuh -- the fused update_stress! + update_velocity! sweep (fused_sv.cuh) with its operand traffic moved off the
// warps' critical path (examples/stokes_3d_inc_ve_T.jl:23-57).
//
// What the ncu source view of k_fused_sv showed (profiles/r2_c1_*): 47 % of all warp samples sit on the FIRST use of phase
// A's 19 global loads -- every warp issues its loads, then waits a loaded DRAM latency (~1.4 us) per plane, and 160 registers
// leave only 12 warps per SM to overlap that wait.  Here
//   * 11 of the 13 operands that are read exactly once (tau[6], tau_old[5]) are brought in by the TMA unit: every warp issues
//     the `cp.async.bulk` row copies (512 B, 16-byte aligned thanks to the PITCHED layout) of ITS row of plane kp+2 into a
//     two-slot shared-memory ring while planes kp and kp+1 are being computed; completion is an mbarrier transaction
//     count, no register and no LSU slot is held while the bytes are in flight;
//   * the three velocity components (re-used between neighbouring rows and planes), Pr and tau_old.yz -- what does not fit
//     the ring next to the exchange buffers at 3 CTAs per SM -- are requested one plane ahead into registers (8 x 128-bit
//     loads in flight across the barrier and phase B);
//   * phase B (velocity update from the freshly computed stresses, shared-memory / DSMEM exchange) is fused_sv.cuh's.
// Arithmetic, order of operations, tile geometry, ping-pong buffers and the barrier protocol are those of fused_sv.cuh;
// the file compiles with nvcc (kernel in ops_fused.cu) and with a host compiler (tests/emul/fused_emul.cpp runs the same
// phase functions thread by thread, the bulk copies as memcpy, and compares bit for bit with the oracle).
#pragma once
#include "fused_sv.cuh"

constexpr int FTM_NS = 11;   // operands staged by the TMA unit: tau xx yy zz xy xz yz, tau_old xx yy zz xy xz

// shared memory of one CTA, in doubles: ring[2 slots][11 operands][TYB rows][64 cells] | fused_sv's exchange buffer | 2 mbarriers per row
FHD int ftm_ring_off(int tyb, int slot, int op, int row) { return ((slot * FTM_NS + op) * tyb + row) * 64; }
FHD int ftm_xch_base(int tyb) { return 2 * FTM_NS * tyb * 64; }
static inline size_t ftm_smem_bytes(int tyb) { return (size_t)(2 * FTM_NS * tyb * 64) * sizeof(double) + fsv_smem_bytes(tyb) + 64; }

struct FusedM {
    FusedT t;                                              // geometry + the carried V planes, as in fused_sv.cuh
    d2 nvx, nvxjm, nvy, nvyjp, nvzkp, nvzjmkp, npr, noyz;  // register-fed operands of the NEXT plane, in flight
};

// ---- the row copies of one plane.  A warp owns one row of the tile and is the only reader of that row's ring entries, so
// every warp feeds itself: one lane issues the 11 copies of its row (two planes ahead) and the warp waits on its own
// mbarrier -- no CTA-wide synchronisation is involved in the ring.  The transport is cp.async.bulk on the device and memcpy
// in the host emulation.  bytes == 0: the row is not addressed (above the box).
struct FtmRow {
    const double* src;
    int dst;        // ring offset (doubles)
    int bytes;
};

// s: the state of ANY lane of the row (its offsets point at plane kp of the sweep); dz: planes ahead
FHD FtmRow ftm_row(const FusedP& p, const FusedT& s, int tyb, int op, int dz, int slot) {
    FtmRow r;
    r.src = nullptr; r.dst = 0; r.bytes = 0;
    if (s.j > p.hi[1]) return r;
    // storage class of the operand: xx yy zz -> CC ; xy -> VV ; xz -> VC ; yz -> CV
    const int comp = op < 6 ? op : op - 6;
    const double* base = op < 6 ? p.tc[comp] : p.to[comp];
    const Strides st = comp < 3 ? p.cc : (comp == 3 ? p.vv : (comp == 4 ? p.vc : p.cv));
    const long long off = comp < 3 ? s.cc : (comp == 3 ? s.vv : (comp == 4 ? s.vc : s.cv));
    const int i0 = s.i - 2 * s.lane;
    // cells [i0, i0 + len): never past the 16-element slack behind the row's pitch (the allocation ends 32 elements
    // behind its last row); i0 is even and the pitch a multiple of 16, so the copy stays 16-byte granular
    int len = st.sy + 14 - i0;
    len = len > 64 ? 64 : len;
    r.src = base + (off - 2 * s.lane) + (long long)dz * st.sz;
    r.dst = ftm_ring_off(tyb, slot, op, s.ty);
    r.bytes = len * 8;
    return r;
}

// part 1: what comes from DRAM (long latency) ; part 2: the neighbouring rows of V (loaded by the neighbouring warps at the
// same time: cache hits, short latency)
FHD void ftm_load_regs(FusedM& m, const FusedP& p, int dz, int part) {
    FusedT& s = m.t;
    if (!s.s_act) return;
    const long long cc = s.cc + (long long)dz * p.cc.sz, vc = s.vc + (long long)dz * p.vc.sz, cv = s.cv + (long long)dz * p.cv.sz;
    if (part & 1) {
        m.nvx   = ld2(p.Vc[0] + vc);
        m.nvy   = ld2(p.Vc[1] + cv);
        m.nvzkp = ld2(p.Vc[2] + cc + p.cc.sz);
        m.npr   = ld2(p.Prc + cc);
        m.noyz  = ld2(p.to[5] + cv);
    }
    if (part & 2) {
        m.nvxjm   = ld2(p.Vc[0] + vc - (long long)s.jm * p.vc.sy);
        m.nvyjp   = ld2(p.Vc[1] + cv + (long long)s.jp * p.cv.sy);
        m.nvzjmkp = ld2(p.Vc[2] + cc - (long long)s.jm * p.cc.sy + p.cc.sz);
    }
}

FHD void ftm_init(FusedM& m, const FusedP& p, int lane, int ty, int grow, int bx, int cyc, int bz, bool fun) {
    fsv_init(m.t, p, lane, ty, grow, bx, cyc, bz, fun);
    const d2 z = fsv_zero();
    m.nvx = z; m.nvxjm = z; m.nvy = z; m.nvyjp = z; m.nvzkp = z; m.nvzjmkp = z; m.npr = z; m.noyz = z;
    ftm_load_regs(m, p, 0, 3);          // plane k0 - 1
}

// ---- phase A: stresses of plane kp from the ring slot + the prefetched registers; stores; requests plane kp+1's registers
template <bool TD>
FHD void ftm_phase_a(FusedM& m, const FusedP& p, int kp, const double* ring, int tyb, int slot, bool more, d2 sn[FSV_NF]) {
    FusedT& s = m.t;
    const d2 z2 = fsv_zero();
    const d2 vx = m.nvx, vxjm = m.nvxjm, vy = m.nvy, vyjp = m.nvyjp, vzkp = m.nvzkp, vzjmkp = m.nvzjmkp, pr = m.npr;
    const d2 oyz = m.noyz;
    if (more) ftm_load_regs(m, p, 1, 1);    // plane kp+1 from DRAM: in flight during phase A, the barrier and phase B
    d2 t[6], o[6];
    const int c2 = 2 * s.lane;
    if (s.s_act) {
#pragma unroll
        for (int c = 0; c < 6; ++c) t[c] = ld2(ring + ftm_ring_off(tyb, slot, c, s.ty) + c2);
#pragma unroll
        for (int c = 0; c < 5; ++c) o[c] = ld2(ring + ftm_ring_off(tyb, slot, 6 + c, s.ty) + c2);
        o[5] = oyz;
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) { t[c] = z2; o[c] = z2; }
    }
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
        const double d3  = div_u<TD>(d, p.three);
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
    if (more) ftm_load_regs(m, p, 1, 2);    // plane kp+1, neighbouring rows: in flight during the barrier and phase B
}
