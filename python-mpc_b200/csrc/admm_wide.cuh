// A latency-oriented variant of the ADMM sweeps: 8 lanes per QP.
//
// After the re-tiling only a few percent of the batch is left (the stragglers), the data is L2-resident and the launch
// is latency-bound: a warp of admm_tma_kernel walks 42 dependent stage sweeps of ~1100 instructions per iteration on
// its own.  Here a QP is spread over a group of 8 lanes — lane a owns component a of the stage vector w = [x | u]
// (a < NX: state row / column a, NX <= a < NW: input a - NX; NW <= 8) — and the 6x6 products of a stage become
// one multiply-add per lane and operand, the operands fetched from the owning lane by a warp shuffle.  A stage sweep
// is ~250 instructions per warp (4 QPs) instead of ~1100, and there are 4x more warps to hide the shuffle latency.
//
// It runs steady-state iterations only: rows in p-form on entry and exit, no termination test — the host loop
// (run_admm) hands the tested iteration to admm_tma_kernel; the iteration before it also leaves its new state in the
// old-state buffer (what the certificates of the tested iteration call x^{k-1}, y^{k-1}).  Every scalar is
// computed by the formulas of qp_thread.cuh in the same order (the kernels agree to the last bits, not bitwise:
// the compiler is free to contract multiply-adds differently).
#pragma once
#include "admm_kernel.cuh"

#if defined(__CUDACC__) && !defined(MPCB_EMU)
namespace mpcb {

constexpr int WIDE_G = 8;      // lanes per QP

template <typename T>
__device__ __forceinline__ T gshfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src, WIDE_G); }

// sum_d c[d] * v_d with v_d fetched from lane d of the group: three independent partial sums (the stage is bound by the
// FP64 dependent-issue latency, a single running sum would be a chain of NW multiply-adds)
template <typename T, int NT>
__device__ __forceinline__ T gdot(const T (&c)[NT], T v) {
    T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int d = 0; d < NT; d += 3) {
        s0 += c[d] * gshfl(v, d);
        if (d + 1 < NT) s1 += c[d + 1] * gshfl(v, d + 1);
        if (d + 2 < NT) s2 += c[d + 2] * gshfl(v, d + 2);
    }
    return (s0 + s1) + s2;
}

// what one lane needs of one stage record
template <typename T, typename L>
struct WideSlice {
    T Lrow[L::NW], Lcol[L::NW];      // row a / column a of Linv_k
    T Da, Edn, pdn, Eb, pb, xv, Dsl, xs, tt, gk;
    T xr, lo, hi;                    // reference and state box of this component at this stage
    T col[L::NX], row[L::NW];        // TV: column a and row a of [A_k | B_k] (else loop-invariant, kept by the kernel)
};
template <typename T, typename L, bool BWD, bool TV>
__device__ __forceinline__ void wide_load(const KParams<T>& p, const T* R, int bb, int k, int a, bool isx, bool isu, int jx,
                                          int ju, bool last, WideSlice<T, L>& s) {
    constexpr int NW = L::NW, NS = L::NS;
    const int aa = a < NW ? a : 0;
#pragma unroll
    for (int d = 0; d < NW; ++d) {
        s.Lrow[d] = (a < NW && d <= a) ? MPCB_AT(R, L::R_F + aa * (aa + 1) / 2 + d) : (T)0;
        s.Lcol[d] = (a < NW && d >= a) ? MPCB_AT(R, L::R_F + d * (d + 1) / 2 + aa) : (T)0;
    }
    s.Da = 1; s.Edn = 1; s.pdn = 0; s.Eb = 1; s.pb = 0; s.xv = 0; s.Dsl = 1; s.xs = 0; s.gk = 0;
    s.xr = 0; s.lo = 0; s.hi = 0;
    if (isx) {
        s.xr = p.Xr[((p.xr_tv ? (size_t)k * L::NX : 0) + jx) * p.ld + bb];
        s.lo = p.xbox ? p.xbox[(k * 2 + 0) * L::NX + jx] : p.xmin[jx];
        s.hi = p.xbox ? p.xbox[(k * 2 + 1) * L::NX + jx] : p.xmax[jx];
    } else if (isu) {
        s.lo = p.umin[ju]; s.hi = p.umax[ju];
    }
    if (TV && !last) {                 // the stage's own linearisation (Ad_list[k], Bd_list[k])
        constexpr int NX = L::NX, NU = L::NU;
        const size_t bo = p.model_bs ? (size_t)bb : 0, ldm = p.model_bs ? p.ld : 1;
        const size_t oa = (size_t)k * NX * NX, ob = (size_t)k * NX * NU;
#pragma unroll
        for (int i = 0; i < NX; ++i)
            s.col[i] = isx ? p.Ad[(oa + i * NX + jx) * ldm + bo] : (isu ? p.Bd[(ob + i * NU + ju) * ldm + bo] : (T)0);
#pragma unroll
        for (int j = 0; j < NW; ++j)
            s.row[j] = !isx ? (T)0 : (j < NX ? p.Ad[(oa + jx * NX + (j < NX ? j : 0)) * ldm + bo]
                                           : p.Bd[(ob + jx * NU + (j >= NX ? j - NX : 0)) * ldm + bo]);
    }
    s.tt = (BWD && a < NW) ? MPCB_AT(R, L::R_T + aa) : (T)0;
    if (isx) {
        s.Da = MPCB_AT(R, L::R_D + L::OX + jx);
        s.Eb = MPCB_AT(R, L::R_E + L::OBX + jx);
        s.pb = MPCB_AT(R, L::R_P + L::OBX + jx);
        s.xv = MPCB_AT(R, L::R_X + L::OX + jx);
        if (NS) { s.Dsl = MPCB_AT(R, L::R_D + L::OS + (NS ? jx : 0)); s.xs = MPCB_AT(R, L::R_X + L::OS + (NS ? jx : 0)); }
        if (!last) {
            s.Edn = MPCB_AT(R, L::R_E + L::ODN + jx);
            s.pdn = MPCB_AT(R, L::R_P + L::ODN + jx);
            s.gk = model_g<T, L>(p, bb, k, jx);
        }
    } else if (isu && !last) {
        s.Da = MPCB_AT(R, L::R_D + L::OU + ju);
        s.Eb = MPCB_AT(R, L::R_E + L::OBU + ju);
        s.pb = MPCB_AT(R, L::R_P + L::OBU + ju);
        s.xv = MPCB_AT(R, L::R_X + L::OU + ju);
    }
}

template <typename T, typename L, bool TV>
__global__ void __launch_bounds__(128) admm_wide_kernel(const __grid_constant__ KParams<T> p) {
    constexpr int NX = L::NX, NU = L::NU, NW = L::NW, NS = L::NS;
    static_assert(NW <= WIDE_G, "one lane per component of [x | u]");
    const int N = p.N;
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) / WIDE_G);      // workspace slot of this group
    const int a = threadIdx.x % WIDE_G;                                            // component owned by this lane
    const bool in_range = gid < p.B;
    const int b = in_range ? gid : p.B - 1;              // out-of-range groups shadow the last QP (shuffles need every lane), never store
    const int bb = p.qp_map ? p.qp_map[b] : b;           // QP index for inputs and outputs
    const bool wr = in_range && p.status[bb] == kUnsolved && a < NW;
    const bool isx = a < NX, isu = a >= NX && a < NW;
    const int ju = isu ? a - NX : 0, jx = isx ? a : 0;
    Ws<T, L> ws(p, b);
    const T c = MPCB_AT(ws.hdr, L::H_C);
    const T rho = clamp_rho(p.rho), rho_eq = (T)kRhoEqOverRhoIneq * rho, sigma = p.sigma, alpha = p.alpha;
    const bool inf_bounds = p.inf_bounds != 0;
    // column a and row a of [A | B]: loop-invariant for a time-invariant model, part of the stage slice otherwise (TV)
    const size_t bo = p.model_bs ? (size_t)bb : 0, ldm = p.model_bs ? p.ld : 1;
    T col0[NX], row0[NW];
#pragma unroll
    for (int i = 0; i < NX; ++i)
        col0[i] = TV ? (T)0 : (isx ? p.Ad[(size_t)(i * NX + jx) * ldm + bo] : (isu ? p.Bd[(size_t)(i * NU + ju) * ldm + bo] : (T)0));
#pragma unroll
    for (int j = 0; j < NW; ++j)
        row0[j] = (TV || !isx) ? (T)0 : (j < NX ? p.Ad[(size_t)(jx * NX + (j < NX ? j : 0)) * ldm + bo]
                                                : p.Bd[(size_t)(jx * NU + (j >= NX ? j - NX : 0)) * ldm + bo]);
#define MPCB_COL(i) (TV ? cur.col[i] : col0[i])
#define MPCB_ROW(j) (TV ? cur.row[j] : row0[j])
    const T xinit = isx ? p.x_init[(size_t)jx * p.ld + bb] : (T)0;
    const T Sj = (NS && isx) ? p.S[jx] : (T)0, Wj = (NS && isx) ? p.W[jx] : (T)0;
    const T E0 = isx ? MPCB_AT(ws.hdr, L::H_E0 + jx) : (T)1;
    const T beq0 = -E0 * xinit;
    T P0 = isx ? MPCB_AT(ws.hdr, L::H_P0 + jx) : (T)0;       // row dyn_0 of this lane, kept in a register across iterations

    WideSlice<T, L> cur, nxt;
    for (int it = p.it0 + 1; it <= p.it_stop; ++it) {
        // ------------------------------------------------------------------ forward sweep (slice k+1 in flight while k is computed)
        T Ed_cur = E0, vd_cur = 0, cprev = 0;
        if (isx) {
            const T z = tmin(tmax(P0, beq0), beq0), yr = P0 - z;
            vd_cur = rho_eq * (z - yr);
        }
        wide_load<T, L, false, TV>(p, ws.R(0), bb, 0, a, isx, isu, jx, ju, N == 0, cur);
        for (int k = 0; k <= N; ++k) {
            const bool last = (k == N);
            if (!last) wide_load<T, L, false, TV>(p, ws.R(k + 1), bb, k + 1, a, isx, isu, jx, ju, k + 1 == N, nxt);
            T* Rw = ws.R(k);
            const T Qj = isx ? (last ? p.QN[jx] : p.Q[jx]) : (T)0;
            const T Da = cur.Da, Ed_next = cur.Edn;
            T vd_next = 0, r = 0;
            if (isx && !last) {
                const T beq = -Ed_next * cur.gk;
                const T z = tmin(tmax(cur.pdn, beq), beq), yr = cur.pdn - z;
                vd_next = rho_eq * (z - yr);
            }
            const T wv = Ed_next * vd_next;
            T colv[NX], rowv[NW];
#pragma unroll
            for (int i = 0; i < NX; ++i) colv[i] = last ? (T)0 : MPCB_COL(i);
#pragma unroll
            for (int j = 0; j < NW; ++j) rowv[j] = last ? (T)0 : MPCB_ROW(j);
            const T acc = gdot<T, NX>(colv, wv);
            if (isx) {
                const T Ebx = cur.Eb;
                const T bx = Ebx * Da, lb = Ebx * cur.lo, ub = Ebx * cur.hi;
                const T rb = row_rho(inf_bounds, lb, ub, rho, rho_eq);
                const T z = tmin(tmax(cur.pb, lb), ub), yr = cur.pb - z;
                const T vbx = rb * (z - yr);
                const T qh = c * Da * (-(Qj * cur.xr));
                const T ex = Ed_cur * Da;
                T v = sigma * cur.xv - qh - ex * vd_cur + bx * vbx + Da * acc;
                if (NS) {
                    const T bs = Sj * Ebx * cur.Dsl;
                    const T mss = c * Wj * cur.Dsl * cur.Dsl + sigma + rb * bs * bs;
                    const T mxs = rb * bx * bs;
                    const T rs = sigma * cur.xs + bs * vbx;
                    v -= mxs * fast_rcp(mss) * rs;
                }
                if (k > 0) v += rho_eq * ex * Ed_cur * cprev;
                r = v;
            } else if (isu && !last) {
                const T bu = cur.Eb * Da, lb = cur.Eb * cur.lo, ub = cur.Eb * cur.hi;
                const T rb = row_rho(inf_bounds, lb, ub, rho, rho_eq);
                const T z = tmin(tmax(cur.pb, lb), ub), yr = cur.pb - z;
                r = sigma * cur.xv + bu * (rb * (z - yr)) + Da * acc;
            }
            // t = Linv r
            const T t = gdot<T, NW>(cur.Lrow, r);       // (Lrow / Lcol are zero outside the triangle)
            if (wr) MPCB_AT(Rw, L::R_T + (a < NW ? a : 0)) = t;
            // g = Linv' t ;  h = D (.) g ;  cprev = [A B] h
            const T g = gdot<T, NW>(cur.Lcol, t);
            const T h = Da * g;
            const T cn = gdot<T, NW>(rowv, h);
            if (!last) cprev = cn;
            Ed_cur = Ed_next; vd_cur = vd_next;
            if (!last) cur = nxt;
        }
        // ------------------------------------------------------------------ backward sweep (slice k-1 in flight while k is computed)
        T xt_next = 0, Dx_next = 1;
        const bool save = wr && admm_is_tested(p, it + 1);       // duplicate the new state into the old-state buffer
        wide_load<T, L, true, TV>(p, ws.R(N), bb, N, a, isx, isu, jx, ju, true, cur);
        for (int k = N; k >= 0; --k) {
            const bool last = (k == N);
            if (k > 0) wide_load<T, L, true, TV>(p, ws.R(k - 1), bb, k - 1, a, isx, isu, jx, ju, false, nxt);
            T* Rw = ws.R(k);
            T* Ow = ws.S(k);
            const T Da = cur.Da, Ed_next = cur.Edn;
            const T exn = Ed_next * Dx_next;             // ex_{k+1} = E_dyn(k+1) D_x(k+1)   (x lanes)
            T rhs = cur.tt;
            const T om = isx ? Ed_next * exn * xt_next : (T)0;
            T colv[NX], rowv[NW];
#pragma unroll
            for (int i = 0; i < NX; ++i) colv[i] = last ? (T)0 : MPCB_COL(i);
#pragma unroll
            for (int j = 0; j < NW; ++j) rowv[j] = last ? (T)0 : MPCB_ROW(j);
            const T acc = gdot<T, NX>(colv, om);
            const T cv = -rho_eq * Da * acc;
            const T sub = gdot<T, NW>(cur.Lrow, cv);
            if (!last) rhs -= sub;
            const T w = gdot<T, NW>(cur.Lcol, rhs);
            // rows dyn_{k+1} need D (.) w of every component
            const T Dw = Da * w;
            const T accd = gdot<T, NW>(rowv, Dw);
            if (isx) {
                const T Ebx = cur.Eb;
                const T bx = Ebx * Da, lb = Ebx * cur.lo, ub = Ebx * cur.hi;
                const T rb = row_rho(inf_bounds, lb, ub, rho, rho_eq);
                Row<T> rw;
                rw.z = tmin(tmax(cur.pb, lb), ub); rw.yr = cur.pb - rw.z;
                T ztil = bx * w;
                if (NS) {
                    const T bs = Sj * Ebx * cur.Dsl;
                    const T mss = c * Wj * cur.Dsl * cur.Dsl + sigma + rb * bs * bs;
                    const T mxs = rb * bx * bs;
                    const T sold = cur.xs;
                    const T rs = sigma * sold + bs * (rb * (rw.z - rw.yr));
                    const T st = (rs - mxs * w) * fast_rcp(mss);
                    ztil += bs * st;
                    const T sn = alpha * st + ((T)1 - alpha) * sold;
                    if (wr) MPCB_AT(Rw, L::R_X + L::OS + (NS ? jx : 0)) = sn;
                    if (save) MPCB_AT(Ow, L::OS + (NS ? jx : 0)) = sn;
                }
                const T pn = row_next(ztil, rw, alpha);
                const T xn = alpha * w + ((T)1 - alpha) * cur.xv;
                if (wr) { MPCB_AT(Rw, L::R_P + L::OBX + jx) = pn; MPCB_AT(Rw, L::R_X + L::OX + jx) = xn; }
                if (save) { MPCB_AT(Ow, L::VS + L::OBX + jx) = pn; MPCB_AT(Ow, L::OX + jx) = xn; }
                if (!last) {
                    // row dyn_{k+1}:  E (A D x~_k + B D u~_k) - ex_{k+1} x~_{k+1} = -E g_k
                    const T zt = Ed_next * accd - exn * xt_next;
                    const T beq = -Ed_next * cur.gk;
                    Row<T> rd;
                    rd.z = tmin(tmax(cur.pdn, beq), beq); rd.yr = cur.pdn - rd.z;
                    const T pdn_new = row_next(zt, rd, alpha);
                    if (wr) MPCB_AT(Rw, L::R_P + L::ODN + jx) = pdn_new;
                    if (save) MPCB_AT(Ow, L::VS + L::ODN + jx) = pdn_new;
                }
                xt_next = w; Dx_next = Da;
            } else if (isu && !last) {
                const T bu = cur.Eb * Da, lb = cur.Eb * cur.lo, ub = cur.Eb * cur.hi;
                Row<T> rw;
                rw.z = tmin(tmax(cur.pb, lb), ub); rw.yr = cur.pb - rw.z;
                const T pn = row_next(bu * w, rw, alpha);
                const T un = alpha * w + ((T)1 - alpha) * cur.xv;
                if (wr) { MPCB_AT(Rw, L::R_P + L::OBU + ju) = pn; MPCB_AT(Rw, L::R_X + L::OU + ju) = un; }
                if (save) { MPCB_AT(Ow, L::VS + L::OBU + ju) = pn; MPCB_AT(Ow, L::OU + ju) = un; }
            }
            if (k > 0) cur = nxt;
        }
        // rows dyn_0 (header)
        if (isx) {
            Row<T> rw;
            rw.z = tmin(tmax(P0, beq0), beq0); rw.yr = P0 - rw.z;
            P0 = row_next(-(E0 * Dx_next) * xt_next, rw, alpha);
            if (save) MPCB_AT(ws.scr_hdr, jx) = P0;
        }
    }
    if (isx && wr) MPCB_AT(ws.hdr, L::H_P0 + jx) = P0;
#undef MPCB_COL
#undef MPCB_ROW
}

}  // namespace mpcb
#endif
