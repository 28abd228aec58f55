// The shared-KKT variant of the ADMM iteration: ONE linearisation for the whole batch (BASELINE configs[1]; the
// single-vehicle closed loop of vehicle_lateral_mpc_slack_increment.py).
//
// When every QP of a batch has the same model, weights and bounds AND the same Ruiz scaling (D, E, c — they depend on
// the QP's own q = -Q xr only through the cost normalisation; equal whenever the scaled ||q||_inf stays below the mean
// column norm of P, e.g. xr = 0), the reduced KKT matrix  M = P^ + sigma I + A^' diag(rho) A^  (slack variables
// eliminated, block tridiagonal over the stages) is ONE matrix for the batch.  Its explicit inverse (n_w x n_w,
// n_w = (N+1)(nx+nu) = 105 for configs[1]; cond(M) ~ 4e3 after Ruiz scaling, so the explicit inverse is accurate to
// ~1e-13) is formed once per setup from the cached block Cholesky factor (dense_inverse_column), and the linear solve of
// an iteration becomes a dense FP64 GEMM over all right-hand sides
//         W (8 QPs x n_w) = R (8 QPs x n_w) . Minv (n_w x n_w)
// on the tensor cores (DMMA, mma.sync.m8n8k4.f64), operands in shared memory.  Everything else of the iteration — the
// right-hand side sigma x - q + A'(rho z - y), the projection on [l, u], the dual step — has no sequential dependency
// left once the solve is dense: lane k of a warp owns stage k of the warp's QP (its x, s, u and row states live in
// registers; the neighbour stage's values come by warp shuffle).
//
// admm_dense_kernel runs the WHOLE loop of such a batch in one launch — iterations, termination tests (residual norms
// reduced over a QP's stages by warp shuffles), infeasibility certificates, exit pass — with no host round trip.
// A CTA = 8 warps = 8 QPs.
#pragma once
#include "admm_kernel.cuh"

namespace mpcb {

constexpr int DENSE_QPB = 8;          // QPs per CTA = M of the DMMA tile
constexpr int DENSE_MAX_NW = 128;     // largest padded n_w whose inverse fits next to the rest in shared memory

MPCB_HD int dense_nwp(int N, int NW) { return ((N + 1) * NW + 7) / 8 * 8; }

// Column `col` of M^-1 (= row `col`: M is symmetric) from the cached factor of QP slot 0:  M = L L', L block lower
// bidiagonal with inverted diagonal blocks Linv_k in the records and coupling blocks F_k = C_k Linv_k' re-applied from
// the model and the scalings (same recurrences as admm_fwd_stage / admm_bwd_stage, right-hand side e_col).
template <typename T, typename L>
MPCB_HD void dense_inverse_column(const KParams<T>& p, int col, T* Minv, int nwp, T* tbuf /* (N+1)*NW scratch */) {
    constexpr int NX = L::NX, NU = L::NU, NW = L::NW;
    const int N = p.N;
    Ws<T, L> ws(p, 0);
    const T rho = clamp_rho(p.rho), rho_eq = (T)kRhoEqOverRhoIneq * rho;
    Model<T, L> m;
    if (!p.tv) load_model<T, L>(p, 0, 0, m);
    T Ed_cur[NX], cprev[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { Ed_cur[i] = MPCB_AT(ws.hdr, L::H_E0 + i); cprev[i] = 0; }
    for (int k = 0; k <= N; ++k) {
        const bool last = (k == N);
        const T* R = ws.R(k);
        if (p.tv && !last) load_model<T, L>(p, 0, k, m);
        T D[NW], Ed_next[NX], r[NW], t[NW], Li[L::LT];
#pragma unroll
        for (int j = 0; j < NX; ++j) { D[j] = MPCB_AT(R, L::R_D + L::OX + j); Ed_next[j] = last ? (T)1 : MPCB_AT(R, L::R_E + L::ODN + j); }
#pragma unroll
        for (int j = 0; j < NU; ++j) D[NX + j] = last ? (T)1 : MPCB_AT(R, L::R_D + L::OU + j);
#pragma unroll
        for (int e = 0; e < L::LT; ++e) Li[e] = MPCB_AT(R, L::R_F + e);
#pragma unroll
        for (int a = 0; a < NW; ++a) r[a] = (k * NW + a == col) ? (T)1 : (T)0;
        if (k > 0) {
#pragma unroll
            for (int j = 0; j < NX; ++j) r[j] += rho_eq * (Ed_cur[j] * D[j]) * Ed_cur[j] * cprev[j];
        }
#pragma unroll
        for (int a = 0; a < NW; ++a) {
            T acc = 0;
#pragma unroll
            for (int d = 0; d <= a; ++d) acc += Li[a * (a + 1) / 2 + d] * r[d];
            t[a] = acc;
            tbuf[k * NW + a] = acc;
        }
        if (!last) {
            T h[NW];
#pragma unroll
            for (int d = 0; d < NW; ++d) {
                T acc = 0;
#pragma unroll
                for (int a = d; a < NW; ++a) acc += Li[a * (a + 1) / 2 + d] * t[a];
                h[d] = D[d] * acc;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                T acc = 0;
#pragma unroll
                for (int j = 0; j < NX; ++j) acc += m.A[i][j] * h[j];
#pragma unroll
                for (int j = 0; j < NU; ++j) acc += m.B[i][j] * h[NX + j];
                cprev[i] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) Ed_cur[i] = Ed_next[i];
    }
    T xt_next[NX], Dx_next[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { xt_next[i] = 0; Dx_next[i] = 1; }
    for (int k = N; k >= 0; --k) {
        const bool last = (k == N);
        const T* R = ws.R(k);
        if (p.tv && !last) load_model<T, L>(p, 0, k, m);
        T D[NW], Ed_next[NX], rhs[NW], w[NW], Li[L::LT];
#pragma unroll
        for (int j = 0; j < NX; ++j) { D[j] = MPCB_AT(R, L::R_D + L::OX + j); Ed_next[j] = last ? (T)1 : MPCB_AT(R, L::R_E + L::ODN + j); }
#pragma unroll
        for (int j = 0; j < NU; ++j) D[NX + j] = last ? (T)1 : MPCB_AT(R, L::R_D + L::OU + j);
#pragma unroll
        for (int e = 0; e < L::LT; ++e) Li[e] = MPCB_AT(R, L::R_F + e);
#pragma unroll
        for (int a = 0; a < NW; ++a) rhs[a] = tbuf[k * NW + a];
        if (!last) {
            T om[NX], cv[NW];
#pragma unroll
            for (int i = 0; i < NX; ++i) om[i] = Ed_next[i] * (Ed_next[i] * Dx_next[i]) * xt_next[i];
#pragma unroll
            for (int a = 0; a < NW; ++a) {
                T acc = 0;
#pragma unroll
                for (int i = 0; i < NX; ++i) acc += (a < NX ? m.A[i][a < NX ? a : 0] : m.B[i][a >= NX ? a - NX : 0]) * om[i];
                cv[a] = -rho_eq * D[a] * acc;
            }
#pragma unroll
            for (int a = 0; a < NW; ++a) {
                T acc = 0;
#pragma unroll
                for (int d = 0; d <= a; ++d) acc += Li[a * (a + 1) / 2 + d] * cv[d];
                rhs[a] -= acc;
            }
        }
#pragma unroll
        for (int d = 0; d < NW; ++d) {
            T acc = 0;
#pragma unroll
            for (int a = d; a < NW; ++a) acc += Li[a * (a + 1) / 2 + d] * rhs[a];
            w[d] = acc;
            Minv[(size_t)col * nwp + k * NW + d] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { xt_next[i] = w[i]; Dx_next[i] = D[i]; }
    }
}

// Is workspace slot b scaled exactly like slot 0?  (D, E of every stage, E of the dyn_0 rows, the cost scaling c.)
template <typename T, typename L>
MPCB_HD bool dense_same_scaling(const KParams<T>& p, int b) {
    Ws<T, L> w0(p, 0), wb(p, b);
    bool same = true;
    for (int k = 0; k <= p.N; ++k)
        for (int e = 0; e < L::VS + L::CS; ++e) same &= MPCB_AT(w0.R(k), L::R_D + e) == MPCB_AT(wb.R(k), L::R_D + e);
    for (int i = 0; i < L::NX; ++i) same &= MPCB_AT(w0.hdr, L::H_E0 + i) == MPCB_AT(wb.hdr, L::H_E0 + i);
    same &= MPCB_AT(w0.hdr, L::H_C) == MPCB_AT(wb.hdr, L::H_C);
    return same;
}

#if defined(__CUDACC__) && !defined(MPCB_EMU)
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// per-stage constants of the batch (identical for every QP by precondition), one column per stage (lane):
//   cst[e][32]:  e in the enum below, vectors indexed j < NX (or NU)
template <typename L>
struct DenseC {
    static constexpr int NX = L::NX, NU = L::NU;
    static constexpr int DX = 0, LB = DX + NX, UB = LB + NX, RB = UB + NX, BX = RB + NX, KB = BX + NX, KS = KB + NX, BS = KS + NX,
                         MX = BS + NX, MI = MX + NX, E2 = MI + NX, EN = E2 + NX, XN = EN + NX, BQ = XN + NX,
                         BU = BQ + NX, LU = BU + NU, UU = LU + NU, RU = UU + NU, AH = RU + NU, BH = AH + NX * NX,
                         COUNT = BH + NX * NU;
};

// reductions over the stages (lanes) of a QP
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ void test_reduce(TestAcc<T>& t) {
    t.pri = warp_max(t.pri); t.dua = warp_max(t.dua); t.nz = warp_max(t.nz); t.nAx = warp_max(t.nAx); t.nq = warp_max(t.nq);
    t.nAty = warp_max(t.nAty); t.nPx = warp_max(t.nPx); t.nEw = warp_max(t.nEw); t.nDv = warp_max(t.nDv);
    t.aup = warp_max(t.aup); t.alo = warp_min(t.alo); t.lhs = warp_sum(t.lhs); t.qv = warp_sum(t.qv);
}

// The termination test of one QP (warp): lane k evaluates stage k with the stage functions of qp_thread.cuh on the QP's
// records in global memory (the caller has just published x and p there), the norms are reduced over the lanes.  Out of
// line: it runs once every check_termination iterations and must not cost the iteration loop registers.
template <typename L>
__device__ __noinline__ bool dense_termination_test(const KParams<double>& p, const AdmmConst<double, L>& qc, const Ws<double, L>& ws,
                                                    const double* cst_k, int bb, int k, bool first, int it, int& status,
                                                    Resid<double>& rs) {
    typedef double T;
    typedef DenseC<L> C;
    constexpr int NX = L::NX;
    const int N = p.N;
    const bool stage = k <= N, last = k >= N;
    const int kk = stage ? k : N;
    const T rho_eq = qc.rho_eq;
    Model<T, L> m;
    load_model<T, L>(p, bb, p.tv ? (last ? N - 1 : kk) : 0, m);
    bool done = false;
    for (int pass = 0; pass < 2 && !done; ++pass) {     // residuals of (x, y), then certificates of (dx, dy)
        const bool cert = pass == 1;
        TestCarry<T, L> cy;
        TestAcc<T> t;
        test_reset(t);
        // rows dyn_k enter stage k's column sums: E and w = y (or dy) of the previous stage's rows, by shuffle
        T wn[NX], en[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            T w = 0;
            const T beq = cst_k[(C::BQ + i) * 32];
            en[i] = cst_k[(C::EN + i) * 32];
            if (!last) {
                w = rho_eq * (MPCB_AT(ws.R(kk), L::R_P + L::ODN + i) - beq);
                if (cert) w -= rho_eq * old_yr(first, MPCB_AT(ws.S(kk), L::VS + L::ODN + i),
                                               first ? MPCB_AT(ws.Y(kk), L::ODN + i) : (T)0, beq, beq, rho_eq);
            }
            wn[i] = w;
        }
        if (k == 0) admm_test_begin<T, L>(p, qc, bb, ws.hdr, cy, t, cert, first, ws.scr_hdr);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const T we = __shfl_up_sync(0xffffffffu, wn[i], 1), ee = __shfl_up_sync(0xffffffffu, en[i], 1);
            if (k > 0) { cy.wd_cur[i] = we; cy.Ed_cur[i] = ee; }
        }
        if (stage)
            admm_test_stage<T, L>(p, qc, m, bb, k, ws.R(k), ws.R(last ? k : k + 1), cy, t, cert, first, ws.Y(k), ws.S(k),
                                  ws.S(last ? k : k + 1));
        test_reduce(t);
        if (!cert) {
            test_to_resid(t, rs);
            if (admm_residual_test<T, L>(p, qc, rs)) { status = kSolved; done = true; }
            else if (!p.certs && it != p.max_iter) break;
        } else {
            Cert<T> ct;
            test_to_cert(t, ct);
            if (admm_decide<T, L>(p, qc, rs, ct, it == p.max_iter, status)) done = true;
        }
    }
    return done;
}

// The whole ADMM loop of a batch that shares one KKT matrix, one launch, no host round trip: iterations it0+1 .. it_stop
// with the termination test (residual norms and, when it fails, the infeasibility certificates) every check_termination
// iterations — the stage functions of qp_thread.cuh evaluated by lane k for stage k on the QP's records in global
// memory (L2), their norms reduced over the lanes by warp shuffles.  A QP that terminates leaves explicit (z, y),
// iteration count, status and residuals behind like every other kernel; the CTA ends when its 8 QPs have.
template <typename L>
__global__ void __launch_bounds__(DENSE_QPB * 32, 1) admm_dense_kernel(const __grid_constant__ KParams<double> p,
                                                                       const double* __restrict__ Minv_g, int nwp) {
    typedef double T;
    typedef DenseC<L> C;
    constexpr int NX = L::NX, NU = L::NU, NW = L::NW, NS = L::NS;
    extern __shared__ __align__(16) unsigned char dsm[];
    const int N = p.N;
    const int lda = nwp + 4;                                   // row pitch: 8 rows x 4 doubles of a DMMA operand hit 16 distinct bank pairs
    T* const Minv = reinterpret_cast<T*>(dsm);                  // [nwp][lda]
    T* const Rs = Minv + (size_t)nwp * lda;                     // [8][lda]  right-hand sides
    T* const Wsol = Rs + DENSE_QPB * lda;                       // [8][lda]  solutions
    T* const cst = Wsol + DENSE_QPB * lda;                      // [C::COUNT][32]
    const int q = threadIdx.x >> 5, k = threadIdx.x & 31;       // QP of the CTA, stage
    const int slot = blockIdx.x * DENSE_QPB + q;
    const bool in_range = slot < p.B;
    const int b = in_range ? slot : p.B - 1;
    const int bb = p.qp_map ? p.qp_map[b] : b;
    const bool stage = k <= N, last = k == N;
    bool open_ = in_range && p.status[bb] == kUnsolved;         // this QP still iterates (warp-uniform)
    const int kk = stage ? k : N;                               // lanes beyond the horizon shadow the last stage, never store

    for (int i = threadIdx.x; i < nwp * nwp; i += blockDim.x) Minv[(i / nwp) * lda + (i % nwp)] = Minv_g[i];
    for (int i = threadIdx.x; i < 2 * DENSE_QPB * lda; i += blockDim.x) Rs[i] = 0;

    Ws<T, L> ws(p, b);
    T* const Rk = ws.R(kk);
    AdmmConst<T, L> qc;
    admm_setup_const<T, L>(p, bb, ws, qc);
    const T c = qc.c, rho = qc.rho, rho_eq = qc.rho_eq, sigma = p.sigma, alpha = p.alpha;
    // ---- per-stage constants (warp 0 fills the table; identical for every QP of the batch)
    if (q == 0) {
        {   // the model of this stage, pre-scaled by the column scalings:  AH[i][j] = A_ij D_x(j),  BH[i][j] = B_ij D_u(j)
            Model<T, L> m;
            load_model<T, L>(p, bb, p.tv ? (last ? N - 1 : kk) : 0, m);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
#pragma unroll
                for (int j = 0; j < NX; ++j)
                    cst[(C::AH + i * NX + j) * 32 + k] = last ? (T)0 : m.A[i][j] * MPCB_AT(Rk, L::R_D + L::OX + j);
#pragma unroll
                for (int j = 0; j < NU; ++j)
                    cst[(C::BH + i * NU + j) * 32 + k] = last ? (T)0 : m.B[i][j] * MPCB_AT(Rk, L::R_D + L::OU + j);
            }
        }
        T lo[NX], hi[NX];
        stage_box<T, L>(p, kk, lo, hi);
        const T* Rn = ws.R(last ? kk : kk + 1);
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const T Dx = MPCB_AT(Rk, L::R_D + L::OX + j), Ebx = MPCB_AT(Rk, L::R_E + L::OBX + j);
            const T En = last ? (T)1 : MPCB_AT(Rk, L::R_E + L::ODN + j);
            const T lb = Ebx * lo[j], ub = Ebx * hi[j];
            const T rb = row_rho(lb, ub, rho, rho_eq);
            const T bx = Ebx * Dx;
            T bs = 0, mxs = 0, mi = 0;
            if (NS) {
                const T Dsl = MPCB_AT(Rk, L::R_D + L::OS + (NS ? j : 0));
                bs = p.S[j] * Ebx * Dsl;
                mi = fast_rcp(c * p.W[j] * Dsl * Dsl + sigma + rb * bs * bs);
                mxs = rb * bx * bs;
            }
            const T beq = last ? (T)0 : -En * model_g<T, L>(p, bb, kk, j);
            cst[(C::DX + j) * 32 + k] = Dx; cst[(C::LB + j) * 32 + k] = lb; cst[(C::UB + j) * 32 + k] = ub;
            cst[(C::RB + j) * 32 + k] = rb; cst[(C::BX + j) * 32 + k] = bx;
            cst[(C::KB + j) * 32 + k] = bx - mxs * mi * bs; cst[(C::KS + j) * 32 + k] = mxs * mi;
            cst[(C::BS + j) * 32 + k] = bs; cst[(C::MX + j) * 32 + k] = mxs; cst[(C::MI + j) * 32 + k] = mi;
            cst[(C::E2 + j) * 32 + k] = last ? (T)0 : En * rho_eq;
            cst[(C::EN + j) * 32 + k] = En;
            cst[(C::XN + j) * 32 + k] = last ? (T)0 : En * MPCB_AT(Rn, L::R_D + L::OX + j);
            cst[(C::BQ + j) * 32 + k] = beq;
        }
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            const T Du = last ? (T)1 : MPCB_AT(Rk, L::R_D + L::OU + j), Ebu = MPCB_AT(Rk, L::R_E + L::OBU + j);
            const T lb = Ebu * p.umin[j], ub = Ebu * p.umax[j];
            cst[(C::BU + j) * 32 + k] = Ebu * Du; cst[(C::LU + j) * 32 + k] = lb; cst[(C::UU + j) * 32 + k] = ub;
            cst[(C::RU + j) * 32 + k] = row_rho(lb, ub, rho, rho_eq);
        }
    }
    // ---- this QP's iterates of stage k (registers), its linear cost term and the dyn_0 rows (lane 0)
    T x[L::VS], pr[L::CS], cqxr[NX];
    T* const P0 = cst + C::COUNT * 32 + q * (3 * NX);            // dyn_0 rows of this QP (lane 0 only): p, beq, E
    T* const beq0 = P0 + NX;
    T* const e0 = beq0 + NX;
    const bool cold = p.it0 == 0 && !p.warm;
#pragma unroll
    for (int e = 0; e < L::VS; ++e) x[e] = cold ? (T)0 : MPCB_AT(Rk, L::R_X + e);
#pragma unroll
    for (int e = 0; e < L::CS; ++e) pr[e] = cold ? (T)0 : MPCB_AT(Rk, L::R_P + e);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const T xr = p.Xr[((p.xr_tv ? (size_t)kk * NX : 0) + j) * p.ld + bb];
        const T* Qk = last ? p.QN : p.Q;
        cqxr[j] = c * MPCB_AT(Rk, L::R_D + L::OX + j) * (Qk[j] * xr);
        if (k == 0) {
            e0[j] = MPCB_AT(ws.hdr, L::H_E0 + j);
            beq0[j] = -e0[j] * p.x_init[(size_t)j * p.ld + bb];
            P0[j] = cold ? (T)0 : MPCB_AT(ws.hdr, L::H_P0 + j);
        }
    }
    if (cold && open_ && stage) {                               // x = z = y = 0 also where the termination test reads them
#pragma unroll
        for (int e = 0; e < L::CS; ++e) MPCB_AT(ws.Y(k), e) = 0;
        if (k == 0) {
#pragma unroll
            for (int i = 0; i < NX; ++i) MPCB_AT(ws.hdr, L::H_Y0 + i) = 0;
        }
    }
    __syncthreads();
#define CST(name, j) cst[(C::name + (j)) * 32 + kk]
    T* const Rq = Rs + q * lda;
    T* const Wq = Wsol + q * lda;
    const int ntile = nwp / 8, ksteps = nwp / 4;
    const int g4 = k >> 2, t4 = k & 3;                          // DMMA fragment coordinates of this lane
    int status = kUnsolved, it_done = p.it_stop;
    Resid<T> rs;
    rs.pri = rs.dua = 0;

    // one iteration; FIRST (iteration 1 of a solve: rows enter as explicit (z, y)) is a compile-time flag, so that the
    // steady-state instantiation carries none of that code.  Returns false when every QP of the CTA has terminated.
    auto iteration = [&](auto first_tag, int it) -> bool {
        constexpr bool first = decltype(first_tag)::value;
        const bool run = open_ && stage;
        // the certificates of an iteration tested at it <= 2 need the state the solve started from / iteration 1 left
        if (run && admm_needs_copy(p, it)) {
            T* O = ws.S(k);
#pragma unroll
            for (int e = 0; e < L::VS; ++e) MPCB_AT(O, e) = x[e];
#pragma unroll
            for (int e = 0; e < L::CS; ++e) MPCB_AT(O, L::VS + e) = pr[e];
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < NX; ++i) MPCB_AT(ws.scr_hdr, i) = P0[i];
            }
        }
        // ------------------------------------------------------------ row states (z, y/rho) of this stage
        T zd[NX], yd[NX], zb[NX], yb[NX], zu[NU], yu[NU], z0[NX], y0[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const Row<T> rd = row_state(first, pr[L::ODN + i], (first && !cold && !last) ? MPCB_AT(ws.Y(kk), L::ODN + i) : (T)0,
                                        CST(BQ, i), CST(BQ, i), qc.rinv_eq());
            zd[i] = rd.z; yd[i] = rd.yr;
            const Row<T> rb_ = row_state(first, pr[L::OBX + i], (first && !cold) ? MPCB_AT(ws.Y(kk), L::OBX + i) : (T)0, CST(LB, i),
                                         CST(UB, i), first ? (T)1 / CST(RB, i) : (T)0);
            zb[i] = rb_.z; yb[i] = rb_.yr;
            z0[i] = 0; y0[i] = 0;
            if (k == 0) {
                const Row<T> r0 = row_state(first, P0[i], (first && !cold) ? MPCB_AT(ws.hdr, L::H_Y0 + i) : (T)0, beq0[i], beq0[i],
                                            qc.rinv_eq());
                z0[i] = r0.z; y0[i] = r0.yr;
            }
        }
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            const Row<T> ru = row_state(first, pr[L::OBU + j], (first && !cold && !last) ? MPCB_AT(ws.Y(kk), L::OBU + j) : (T)0,
                                        CST(LU, j), CST(UU, j), first ? (T)1 / CST(RU, j) : (T)0);
            zu[j] = ru.z; yu[j] = ru.yr;
        }
        // ------------------------------------------------------------ right-hand side  sigma x - q + A'(rho z - y)
        T wv[NX], vb[NX], wprev[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) wv[i] = CST(E2, i) * (zd[i] - yd[i]);                    // E rho_eq (z - y/rho) of rows dyn_{k+1}
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const T up = __shfl_up_sync(0xffffffffu, wv[i], 1);
            wprev[i] = up;                                                                      // rows dyn_k ...
            if (k == 0) wprev[i] = e0[i] * rho_eq * (z0[i] - y0[i]);                            // ... dyn_0: the header
        }
        if (run) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                vb[j] = CST(RB, j) * (zb[j] - yb[j]);
                T acc = -wprev[j] * CST(DX, j);
#pragma unroll
                for (int i = 0; i < NX; ++i) acc += CST(AH, i * NX + j) * wv[i];
                T v = sigma * x[L::OX + j] + cqxr[j] + acc + CST(KB, j) * vb[j];
                if (NS) v -= CST(KS, j) * (sigma * x[L::OS + (NS ? j : 0)]);
                Rq[k * NW + j] = v;
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                T v = 0;
                if (!last) {
                    T acc = 0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) acc += CST(BH, i * NU + j) * wv[i];
                    v = sigma * x[L::OU + j] + CST(BU, j) * (CST(RU, j) * (zu[j] - yu[j])) + acc;
                }
                Rq[k * NW + NX + j] = v;
            }
        }
        // (the barrier that publishes the right-hand sides also tells whether any QP of the CTA still iterates)
        if (!__syncthreads_or(open_)) return false;
        // ------------------------------------------------------------ W = R . Minv on the FP64 tensor cores
        // warp q owns the 8-column tiles q and q + 8 of W: the A fragments (rows of R) serve both, four accumulator pairs
        // give the tensor pipe independent chains
        {
            const bool two = q + DENSE_QPB < ntile;
            const T* Ap = Rs + g4 * lda + t4;                           // A fragment: row = QP g4, col = k-index t4
            const T* Bp = Minv + (size_t)t4 * lda + q * 8 + g4;          // B fragment: row = k-index t4, col = g4 of tile q
            const int stepB = 4 * lda;
            T c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0_ = 0, e1_ = 0, f0 = 0, f1 = 0;
            if (q < ntile) {
#pragma unroll 2
                for (int ks = 0; ks < ksteps; ks += 2) {                 // (ksteps is even: nwp is a multiple of 8)
                    const T a0 = Ap[0], a1 = Ap[4];
                    const T b0 = Bp[0], b1 = Bp[stepB];
                    dmma_m8n8k4(c0, c1, a0, b0);
                    dmma_m8n8k4(d0, d1, a1, b1);
                    if (two) {
                        const T g0 = Bp[8 * DENSE_QPB], g1 = Bp[stepB + 8 * DENSE_QPB];
                        dmma_m8n8k4(e0_, e1_, a0, g0);
                        dmma_m8n8k4(f0, f1, a1, g1);
                    }
                    Ap += 8; Bp += 2 * stepB;
                }
                T* Wp = Wsol + g4 * lda + q * 8 + 2 * t4;               // C fragment: row = QP g4, cols 2 t4, 2 t4 + 1
                Wp[0] = c0 + d0; Wp[1] = c1 + d1;
                if (two) { Wp[8 * DENSE_QPB] = e0_ + f0; Wp[8 * DENSE_QPB + 1] = e1_ + f1; }
            }
        }
        __syncthreads();
        // ------------------------------------------------------------ relaxation, projection, dual step (p-form)
        if (run) {
            T w[NW], xtn[NX];
#pragma unroll
            for (int a = 0; a < NW; ++a) w[a] = Wq[k * NW + a];
#pragma unroll
            for (int i = 0; i < NX; ++i) xtn[i] = last ? (T)0 : Wq[(k + 1) * NW + i];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T zt = CST(BX, j) * w[j];
                if (NS) {
                    const T sold = x[L::OS + (NS ? j : 0)];
                    const T st = (sigma * sold + CST(BS, j) * vb[j] - CST(MX, j) * w[j]) * CST(MI, j);
                    zt += CST(BS, j) * st;
                    x[L::OS + (NS ? j : 0)] = alpha * st + ((T)1 - alpha) * sold;
                }
                pr[L::OBX + j] = alpha * zt + ((T)1 - alpha) * zb[j] + yb[j];
                x[L::OX + j] = alpha * w[j] + ((T)1 - alpha) * x[L::OX + j];
            }
            if (!last) {
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    pr[L::OBU + j] = alpha * (CST(BU, j) * w[NX + j]) + ((T)1 - alpha) * zu[j] + yu[j];
                    x[L::OU + j] = alpha * w[NX + j] + ((T)1 - alpha) * x[L::OU + j];
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    T acc = 0;
#pragma unroll
                    for (int j = 0; j < NX; ++j) acc += CST(AH, i * NX + j) * w[j];
#pragma unroll
                    for (int j = 0; j < NU; ++j) acc += CST(BH, i * NU + j) * w[NX + j];
                    const T zt = CST(EN, i) * acc - CST(XN, i) * xtn[i];
                    pr[L::ODN + i] = alpha * zt + ((T)1 - alpha) * zd[i] + yd[i];
                }
            }
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < NX; ++i) P0[i] = alpha * (-(e0[i] * CST(DX, i)) * w[i]) + ((T)1 - alpha) * z0[i] + y0[i];
            }
            // the next iteration is tested: its certificates need this state as (x^{k-1}, y^{k-1})
            if (admm_saves(p, it)) {
                T* O = ws.S(k);
#pragma unroll
                for (int e = 0; e < L::VS; ++e) MPCB_AT(O, e) = x[e];
#pragma unroll
                for (int e = 0; e < L::CS; ++e) MPCB_AT(O, L::VS + e) = pr[e];
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) MPCB_AT(ws.scr_hdr, i) = P0[i];
                }
            }
        }
        // ------------------------------------------------------------ termination test (auxil.c: check_termination)
        if (open_ && admm_is_tested(p, it)) {                   // warp-uniform
            if (stage) {                                        // the stage functions read the records: publish the new state
                T* Rw = ws.R(k);
#pragma unroll
                for (int e = 0; e < L::VS; ++e) MPCB_AT(Rw, L::R_X + e) = x[e];
#pragma unroll
                for (int e = 0; e < L::CS; ++e) MPCB_AT(Rw, L::R_P + e) = pr[e];
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) MPCB_AT(ws.hdr, L::H_P0 + i) = P0[i];
                }
            }
            __syncwarp();
            const bool done = dense_termination_test<L>(p, qc, ws, cst + kk, bb, k, first, it, status, rs);
            if (done) { open_ = false; it_done = it; }
        }
            return true;
    };
    for (int it = p.it0 + 1; it <= p.it_stop; ++it) {
        const bool go = (it == 1) ? iteration(std::true_type{}, it) : iteration(std::false_type{}, it);
        if (!go) break;
    }
#undef CST
    // ---- leave the state behind: explicit (z, y) for a terminated QP, p-form for one the launch leaves unsolved
    if (in_range && p.status[bb] == kUnsolved && stage) {
        T* Rw = ws.R(k);
#pragma unroll
        for (int e = 0; e < L::VS; ++e) MPCB_AT(Rw, L::R_X + e) = x[e];
#pragma unroll
        for (int e = 0; e < L::CS; ++e) MPCB_AT(Rw, L::R_P + e) = pr[e];
        if (k == 0) {
#pragma unroll
            for (int i = 0; i < NX; ++i) MPCB_AT(ws.hdr, L::H_P0 + i) = P0[i];
        }
        if (status != kUnsolved) {
            admm_exit_stage<T, L>(p, qc, bb, k, Rw, ws.Y(k));
            if (k == 0) {
                admm_exit_header<T, L>(p, qc, bb, ws.hdr);
                p.iter[bb] = it_done; p.pri_res[bb] = rs.pri; p.dua_res[bb] = rs.dua;
            }
        }
    }
    __syncwarp();
    if (in_range && k == 0 && status != kUnsolved && p.status[bb] == kUnsolved) p.status[bb] = status;
}

inline size_t dense_smem_bytes(int nwp, int ncst) {
    const size_t lda = (size_t)nwp + 4;
    return ((size_t)nwp * lda + 2 * DENSE_QPB * lda + (size_t)ncst * 32 + DENSE_QPB * 3 * MAXNX) * sizeof(double);
}
#endif   // __CUDACC__ && !MPCB_EMU

}  // namespace mpcb
