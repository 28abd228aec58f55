// Per-QP device functions of the batched MPC-QP solver: one GPU thread owns one QP.
//
//   scale_one   Ruiz equilibration + cost scaling of the stage-structured KKT matrix
//               (OSQP scaling.c: scale_data; paper Algorithm 2)           -> D, E, c
//   factor_one  cached factorisation of the reduced KKT matrix
//               M = P^ + sigma I + A^' diag(rho) A^   (block tridiagonal over stages after the
//               slack variables are eliminated in closed form) as a block-bidiagonal Cholesky;
//               stores inverse diagonal blocks Linv_k and coupling blocks F_k   -> fac
//   admm_one    the OSQP ADMM loop (osqp.c: update_xz_tilde/update_x/update_z/update_y,
//               auxil.c: check_termination) with fixed rho / sigma / alpha
//
// Replaces, for a batch, the per-QP `osqp.OSQP().setup(P, q, A, l, u); prob.solve()` of
//   /root/reference/Control/MPC/mpc_kinematics.py:205-211
//   /root/reference/Control/MPC/mpc_dynamics.py:248-252, 398-402
//   /root/reference/vehicle_lateral_mpc_slack_increment.py:118-122, 236-250
// The QP itself (P, q, A, l, u of those files) is never materialised on this path: the
// kernels read the stage data (A_k, B_k, g_k, x_init, Xr, weights, bounds) and apply the
// structured operators directly.  qp_build.cuh materialises it for inspection/parity.
//
// All functions are __host__ __device__ so that tests/emu can run the identical code on the
// CPU of the (GPU-less) build container; the product only ever launches them as kernels.
#pragma once
#include "mpc_common.h"

namespace mpcb {

MPCB_HD float msqrt(float v) { return sqrtf(v); }
MPCB_HD double msqrt(double v) { return sqrt(v); }
template <typename T> MPCB_HD T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> MPCB_HD T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> MPCB_HD T tabs(T a) { return a < (T)0 ? -a : a; }
template <typename T> MPCB_HD T limit_scaling(T v) {
    v = v < (T)kMinScaling ? (T)1 : v;
    return v > (T)kMaxScaling ? (T)kMaxScaling : v;
}
template <typename T> MPCB_HD T clip_infty(T v) {
    return tmin(tmax(v, (T)-kOsqpInfty), (T)kOsqpInfty);
}
// auxil.c: set_rho_vec, evaluated on the scaled bounds of one row
template <typename T> MPCB_HD T row_rho(T l, T u, T rho, T rho_eq) {
    if (l < (T)(-kOsqpInfty * kMinScaling) && u > (T)(kOsqpInfty * kMinScaling)) return (T)kRhoMin;
    if (u - l < (T)kRhoTol) return rho_eq;
    return rho;
}
template <typename T> MPCB_HD T clamp_rho(T rho) {
    return tmin(tmax(rho, (T)kRhoMin), (T)kRhoMax);
}

template <typename T, typename L>
struct Model {
    T A[L::NX][L::NX];
    T B[L::NX][L::NU];
    T g[L::NX];
};

template <typename T, typename L>
MPCB_HD void load_model(const KParams<T>& p, int b, int k, Model<T, L>& m) {
    constexpr int NX = L::NX, NU = L::NU;
    const size_t bo = p.model_bs ? (size_t)b : 0;
    const size_t ld = p.model_bs ? p.ld : 1;
    const size_t oa = p.tv ? (size_t)k * NX * NX : 0, ob = p.tv ? (size_t)k * NX * NU : 0,
                 og = p.tv ? (size_t)k * NX : 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) m.A[i][j] = p.Ad[(oa + i * NX + j) * ld + bo];
#pragma unroll
        for (int j = 0; j < NU; ++j) m.B[i][j] = p.Bd[(ob + i * NU + j) * ld + bo];
        m.g[i] = p.gd ? p.gd[(og + i) * ld + bo] : (T)0;
    }
}

template <typename T, typename L>
MPCB_HD void stage_box(const KParams<T>& p, int k, T* lo, T* hi) {
#pragma unroll
    for (int i = 0; i < L::NX; ++i) {
        T a = p.xbox ? p.xbox[(k * 2 + 0) * L::NX + i] : p.xmin[i];
        T c = p.xbox ? p.xbox[(k * 2 + 1) * L::NX + i] : p.xmax[i];
        lo[i] = clip_infty(a);
        hi[i] = clip_infty(c);
    }
}

// ------------------------------------------------------------------------------------------
// Ruiz equilibration (scaling.c: scale_data).  The scaled matrices are never stored: entry
// (i,j) of the scaled A is E_i * A_ij * D_j with the running D, E.
// ------------------------------------------------------------------------------------------
template <typename T, typename L>
MPCB_HD void scale_one(const KParams<T>& p, int b) {
    constexpr int NX = L::NX, NU = L::NU, NS = L::NS;
    const size_t ld = p.ld;
    const int N = p.N;
    const size_t halfD = (size_t)(N + 1) * L::VS * ld, halfE = (size_t)(N + 1) * L::CS * ld;
    T* Dg = p.D + b;
    T* Eg = p.E + b;
    for (int e = 0; e < (N + 1) * L::VS; ++e) Dg[(size_t)e * ld] = (T)1;
    for (int e = 0; e < (N + 1) * L::CS; ++e) Eg[(size_t)e * ld] = (T)1;
    T c = (T)1;
    const T nvar = (T)L::nvar(N);
    Model<T, L> m;
    if (!p.tv) load_model<T, L>(p, b, 0, m);
    for (int it = 0; it < p.scaling; ++it) {
        const T* Ds = Dg + (size_t)(it & 1) * halfD;
        T* Dd = Dg + (size_t)((it & 1) ^ 1) * halfD;
        const T* Es = Eg + (size_t)(it & 1) * halfE;
        T* Ed = Eg + (size_t)((it & 1) ^ 1) * halfE;
        T sumP = 0, maxq = 0;
        T Ed_cur[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) Ed_cur[i] = Es[(size_t)(L::OD + i) * ld];
        for (int k = 0; k <= N; ++k) {
            const bool last = (k == N);
            const size_t vb = (size_t)k * L::VS, cb = (size_t)k * L::CS;
            if (p.tv && !last) load_model<T, L>(p, b, k, m);
            T Dx[NX], Dsl[NX > 0 ? NX : 1], Du[NU], Ebx[NX], Ebu[NU], Ed_next[NX], Dx_next[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                Dx[i] = Ds[(vb + L::OX + i) * ld];
                Dsl[i] = NS ? Ds[(vb + L::OS + i) * ld] : (T)1;
                Ebx[i] = Es[(cb + L::OBX + i) * ld];
                Ed_next[i] = last ? (T)1 : Es[(cb + L::CS + L::OD + i) * ld];
                Dx_next[i] = last ? (T)1 : Ds[(vb + L::VS + L::OX + i) * ld];
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                Du[j] = Ds[(vb + L::OU + j) * ld];
                Ebu[j] = Es[(cb + L::OBU + j) * ld];
            }
            const T* Qk = last ? p.QN : p.Q;
            // ---- column norms of the KKT matrix -> new D
            T Dxn[NX], Dsn[NX], Dun[NU];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T v = c * tabs(Qk[j]) * Dx[j] * Dx[j];
                v = tmax(v, Ed_cur[j] * Dx[j]);
                v = tmax(v, Ebx[j] * Dx[j]);
                if (!last) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) v = tmax(v, tabs(m.A[i][j]) * Ed_next[i] * Dx[j]);
                }
                Dxn[j] = Dx[j] * ((T)1 / msqrt(limit_scaling(v)));
                if (NS) {
                    T w = tmax(c * tabs(p.W[j]) * Dsl[j] * Dsl[j], tabs(p.S[j]) * Ebx[j] * Dsl[j]);
                    Dsn[j] = Dsl[j] * ((T)1 / msqrt(limit_scaling(w)));
                } else {
                    Dsn[j] = (T)1;
                }
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                T v = tmax(c * tabs(p.R[j]) * Du[j] * Du[j], Ebu[j] * Du[j]);
#pragma unroll
                for (int i = 0; i < NX; ++i) v = tmax(v, tabs(m.B[i][j]) * Ed_next[i] * Du[j]);
                Dun[j] = last ? (T)1 : Du[j] * ((T)1 / msqrt(limit_scaling(v)));
            }
            // ---- row norms of A -> new E
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    T v = Ed_cur[i] * Dx[i];
                    Ed[(size_t)(L::OD + i) * ld] = Ed_cur[i] * ((T)1 / msqrt(limit_scaling(v)));
                }
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                if (!last) {
                    T v = Ed_next[i] * Dx_next[i];
#pragma unroll
                    for (int j = 0; j < NX; ++j) v = tmax(v, tabs(m.A[i][j]) * Ed_next[i] * Dx[j]);
#pragma unroll
                    for (int j = 0; j < NU; ++j) v = tmax(v, tabs(m.B[i][j]) * Ed_next[i] * Du[j]);
                    Ed[(cb + L::CS + L::OD + i) * ld] = Ed_next[i] * ((T)1 / msqrt(limit_scaling(v)));
                }
                T w = Ebx[i] * Dx[i];
                if (NS) w = tmax(w, tabs(p.S[i]) * Ebx[i] * Dsl[i]);
                Ed[(cb + L::OBX + i) * ld] = Ebx[i] * ((T)1 / msqrt(limit_scaling(w)));
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                T w = Ebu[j] * Du[j];
                Ed[(cb + L::OBU + j) * ld] = last ? (T)1 : Ebu[j] * ((T)1 / msqrt(limit_scaling(w)));
            }
            // ---- store new D, accumulate the cost-normalisation statistics with it
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                Dd[(vb + L::OX + j) * ld] = Dxn[j];
                sumP += c * tabs(Qk[j]) * Dxn[j] * Dxn[j];
                const T xr = p.Xr[((p.xr_tv ? (size_t)k * NX : 0) + j) * ld + b];
                maxq = tmax(maxq, tabs(c * Dxn[j] * (-(Qk[j] * xr))));
                if (NS) {
                    Dd[(vb + L::OS + j) * ld] = Dsn[j];
                    sumP += c * tabs(p.W[j]) * Dsn[j] * Dsn[j];
                }
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                Dd[(vb + L::OU + j) * ld] = Dun[j];
                if (!last) sumP += c * tabs(p.R[j]) * Dun[j] * Dun[j];
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) Ed_cur[i] = Ed_next[i];
        }
        T c_temp = sumP / nvar;
        const T nq = limit_scaling(maxq);
        c_temp = tmax(c_temp, nq);
        c_temp = (T)1 / limit_scaling(c_temp);
        c *= c_temp;
    }
    if (p.scaling & 1) {
        for (int e = 0; e < (N + 1) * L::VS; ++e) Dg[(size_t)e * ld] = Dg[halfD + (size_t)e * ld];
        for (int e = 0; e < (N + 1) * L::CS; ++e) Eg[(size_t)e * ld] = Eg[halfE + (size_t)e * ld];
    }
    p.c[b] = c;
}

// Scaled per-stage coefficients shared by factor_one / admm_one / residuals.
template <typename T, typename L>
struct StageCoef {
    T ex[L::NX];             // |-I| entry of dyn_k:     E^d_k * D^x_k
    T bx[L::NX];             // bound row on x_k:        E^bx_k * D^x_k
    T bs[L::NX];             // bound row on s_k:        S * E^bx_k * D^s_k
    T bu[L::NU];             // bound row on u_k:        E^bu_k * D^u_k
    T px[L::NX], ps[L::NX], pu[L::NU];      // scaled diagonal of P
    T lbx[L::NX], ubx[L::NX], lbu[L::NU], ubu[L::NU];   // scaled bounds
    T rbx[L::NX], rbu[L::NU];               // rho of the bound rows
    T mss[L::NX], mxs[L::NX];               // slack elimination (M_ss, M_xs)
    T Ah[L::NX][L::NX], Bh[L::NX][L::NU];   // scaled dynamics blocks of row dyn_{k+1}
    T Ed_next[L::NX];                       // E of dyn_{k+1}
};

template <typename T, typename L>
MPCB_HD void stage_coef(const KParams<T>& p, int b, int k, const Model<T, L>& m, const T* Ed_cur, T c,
                        T rho, T rho_eq, StageCoef<T, L>& s) {
    constexpr int NX = L::NX, NU = L::NU, NS = L::NS;
    const size_t ld = p.ld;
    const bool last = (k == p.N);
    const size_t vb = (size_t)k * L::VS, cb = (size_t)k * L::CS;
    const T* Dg = p.D + b;
    const T* Eg = p.E + b;
    T lo[NX], hi[NX];
    stage_box<T, L>(p, k, lo, hi);
    const T* Qk = last ? p.QN : p.Q;
    T Dx[NX], Du[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        Dx[i] = Dg[(vb + L::OX + i) * ld];
        const T Ebx = Eg[(cb + L::OBX + i) * ld];
        s.ex[i] = Ed_cur[i] * Dx[i];
        s.bx[i] = Ebx * Dx[i];
        s.px[i] = c * Qk[i] * Dx[i] * Dx[i];
        s.lbx[i] = Ebx * lo[i];
        s.ubx[i] = Ebx * hi[i];
        s.rbx[i] = row_rho(s.lbx[i], s.ubx[i], rho, rho_eq);
        if (NS) {
            const T Dsl = Dg[(vb + L::OS + i) * ld];
            s.bs[i] = p.S[i] * Ebx * Dsl;
            s.ps[i] = c * p.W[i] * Dsl * Dsl;
            s.mss[i] = s.ps[i] + p.sigma + s.rbx[i] * s.bs[i] * s.bs[i];
            s.mxs[i] = s.rbx[i] * s.bx[i] * s.bs[i];
        } else {
            s.bs[i] = 0; s.ps[i] = 0; s.mss[i] = 1; s.mxs[i] = 0;
        }
        s.Ed_next[i] = last ? (T)1 : Eg[(cb + L::CS + L::OD + i) * ld];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        Du[j] = Dg[(vb + L::OU + j) * ld];
        const T Ebu = Eg[(cb + L::OBU + j) * ld];
        s.bu[j] = Ebu * Du[j];
        s.pu[j] = c * p.R[j] * Du[j] * Du[j];
        s.lbu[j] = Ebu * clip_infty(p.umin[j]);
        s.ubu[j] = Ebu * clip_infty(p.umax[j]);
        s.rbu[j] = row_rho(s.lbu[j], s.ubu[j], rho, rho_eq);
    }
    if (!last) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) s.Ah[i][j] = s.Ed_next[i] * m.A[i][j] * Dx[j];
#pragma unroll
            for (int j = 0; j < NU; ++j) s.Bh[i][j] = s.Ed_next[i] * m.B[i][j] * Du[j];
        }
    } else {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) s.Ah[i][j] = 0;
#pragma unroll
            for (int j = 0; j < NU; ++j) s.Bh[i][j] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Cached factorisation of the reduced KKT matrix (the "cached KKT Cholesky").
// ------------------------------------------------------------------------------------------
template <typename T, typename L>
MPCB_HD void factor_one(const KParams<T>& p, int b) {
    constexpr int NX = L::NX, NU = L::NU, NW = L::NW;
    const size_t ld = p.ld;
    const int N = p.N;
    const T c = p.c[b];
    const T rho = clamp_rho(p.rho), rho_eq = (T)kRhoEqOverRhoIneq * rho;
    Model<T, L> m;
    if (!p.tv) load_model<T, L>(p, b, 0, m);
    T Fprev[NX][NW];
    T Ed_cur[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) Ed_cur[i] = p.E[(size_t)(L::OD + i) * ld + b];
    T* fg = p.fac + b;
    int bad = 0;
    StageCoef<T, L> s;
    for (int k = 0; k <= N; ++k) {
        const bool last = (k == N);
        if (p.tv && !last) load_model<T, L>(p, b, k, m);
        stage_coef<T, L>(p, b, k, m, Ed_cur, c, rho, rho_eq, s);
        T Sm[NW][NW];
#pragma unroll
        for (int a = 0; a < NW; ++a)
#pragma unroll
            for (int d = 0; d < NW; ++d) Sm[a][d] = 0;
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            T d = s.px[j] + p.sigma + rho_eq * s.ex[j] * s.ex[j] + s.rbx[j] * s.bx[j] * s.bx[j];
            if (L::SLACK) d -= s.mxs[j] * s.mxs[j] / s.mss[j];
            Sm[j][j] = d;
        }
#pragma unroll
        for (int j = 0; j < NU; ++j)
            Sm[NX + j][NX + j] = last ? (T)1 : s.pu[j] + p.sigma + s.rbu[j] * s.bu[j] * s.bu[j];
        if (!last) {
#pragma unroll
            for (int a = 0; a < NW; ++a)
#pragma unroll
                for (int d = 0; d <= a; ++d) {
                    T acc = 0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T ga = a < NX ? s.Ah[i][a] : s.Bh[i][a - NX];
                        const T gd = d < NX ? s.Ah[i][d] : s.Bh[i][d - NX];
                        acc += ga * gd;
                    }
                    Sm[a][d] += rho_eq * acc;
                }
        }
        if (k > 0) {
#pragma unroll
            for (int a = 0; a < NX; ++a)
#pragma unroll
                for (int d = 0; d <= a; ++d) {
                    T acc = 0;
#pragma unroll
                    for (int e = 0; e < NW; ++e) acc += Fprev[a][e] * Fprev[d][e];
                    Sm[a][d] -= acc;
                }
        }
        // Cholesky, lower, in place
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            T d = Sm[j][j];
#pragma unroll
            for (int e = 0; e < j; ++e) d -= Sm[j][e] * Sm[j][e];
            if (!(d > (T)0)) { bad = 1; d = (T)1e-30; }
            d = msqrt(d);
            Sm[j][j] = d;
            const T inv = (T)1 / d;
#pragma unroll
            for (int i = j + 1; i < NW; ++i) {
                T v = Sm[i][j];
#pragma unroll
                for (int e = 0; e < j; ++e) v -= Sm[i][e] * Sm[j][e];
                Sm[i][j] = v * inv;
            }
        }
        // inverse of the lower-triangular factor
        T Li[NW][NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            Li[j][j] = (T)1 / Sm[j][j];
#pragma unroll
            for (int i = j + 1; i < NW; ++i) {
                T v = 0;
#pragma unroll
                for (int e = j; e < i; ++e) v += Sm[i][e] * Li[e][j];
                Li[i][j] = -v / Sm[i][i];
            }
        }
        const size_t fb = (size_t)k * L::FAC;
#pragma unroll
        for (int a = 0; a < NW; ++a)
#pragma unroll
            for (int d = 0; d <= a; ++d) fg[(fb + a * (a + 1) / 2 + d) * ld] = Li[a][d];
        if (!last) {
            // coupling block C_k = M[x_{k+1}, w_k] = -rho_eq * ex_{k+1} * [Ah Bh];  F_k = C_k Linv_k'
            T exn[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i)
                exn[i] = s.Ed_next[i] * p.D[((size_t)(k + 1) * L::VS + L::OX + i) * ld + b];
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int a = 0; a < NW; ++a) {
                    T acc = 0;
#pragma unroll
                    for (int d = 0; d <= a; ++d) {
                        const T g = d < NX ? s.Ah[i][d] : s.Bh[i][d - NX];
                        acc += g * Li[a][d];
                    }
                    Fprev[i][a] = -rho_eq * exn[i] * acc;
                    fg[(fb + L::LT + i * NW + a) * ld] = Fprev[i][a];
                }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) Ed_cur[i] = s.Ed_next[i];
    }
    if (bad) p.status[b] = -7;   // OSQP_NON_CVX: reduced KKT matrix lost positive definiteness
}

// Row update of one constraint (osqp.c: update_z, update_y):
//   zr = alpha z~ + (1-alpha) z;  z+ = clip(zr + y/rho);  y+ = y + rho (zr - z+)
template <typename T>
MPCB_HD void row_update(T zt, T l, T u, T rho, T alpha, T& z, T& y) {
    const T zr = alpha * zt + ((T)1 - alpha) * z;
    T zn = zr + y / rho;
    zn = tmin(tmax(zn, l), u);
    y = y + rho * (zr - zn);
    z = zn;
}

// ------------------------------------------------------------------------------------------
// The ADMM loop.
// ------------------------------------------------------------------------------------------
template <typename T, typename L>
MPCB_HD void admm_one(const KParams<T>& p, int b) {
    constexpr int NX = L::NX, NU = L::NU, NW = L::NW, NS = L::NS;
    const size_t ld = p.ld;
    const int N = p.N;
    if (p.status[b] == -7) { p.iter[b] = 0; return; }
    const T c = p.c[b], cinv = (T)1 / c;
    const T rho = clamp_rho(p.rho), rho_eq = (T)kRhoEqOverRhoIneq * rho;
    const T sigma = p.sigma, alpha = p.alpha;
    Model<T, L> m;
    if (!p.tv) load_model<T, L>(p, b, 0, m);
    T* xg = p.x + b;
    T* zg = p.z + b;
    T* yg = p.y + b;
    T* tg = p.t + b;
    const T* fg = p.fac + b;
    const T* Eg = p.E + b;
    const T* Dg = p.D + b;
    if (!p.warm) {
        for (int e = 0; e < (N + 1) * L::VS; ++e) xg[(size_t)e * ld] = 0;
        for (int e = 0; e < (N + 1) * L::CS; ++e) { zg[(size_t)e * ld] = 0; yg[(size_t)e * ld] = 0; }
    }
    T xinit[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xinit[i] = p.x_init[(size_t)i * ld + b];

    int status = kUnsolved, it = 0, checked = 0;
    T pri = 0, dua = 0;
    StageCoef<T, L> s;
    for (it = 1; it <= p.max_iter; ++it) {
        // ================= forward sweep: right-hand side + L^{-1}
        {
            T Ed_cur[NX], vd_cur[NX], tprev[NW];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                Ed_cur[i] = Eg[(size_t)(L::OD + i) * ld];
                vd_cur[i] = rho_eq * zg[(size_t)(L::OD + i) * ld] - yg[(size_t)(L::OD + i) * ld];
            }
#pragma unroll
            for (int a = 0; a < NW; ++a) tprev[a] = 0;
            for (int k = 0; k <= N; ++k) {
                const bool last = (k == N);
                const size_t vb = (size_t)k * L::VS, cb = (size_t)k * L::CS, fb = (size_t)k * L::FAC;
                if (p.tv && !last) load_model<T, L>(p, b, k, m);
                stage_coef<T, L>(p, b, k, m, Ed_cur, c, rho, rho_eq, s);
                const T* Qk = last ? p.QN : p.Q;
                T vd_next[NX], r[NW];
#pragma unroll
                for (int i = 0; i < NX; ++i)
                    vd_next[i] = last ? (T)0
                                      : rho_eq * zg[(cb + L::CS + L::OD + i) * ld] - yg[(cb + L::CS + L::OD + i) * ld];
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    const T xr = p.Xr[((p.xr_tv ? (size_t)k * NX : 0) + j) * ld + b];
                    const T qh = c * Dg[(vb + L::OX + j) * ld] * (-(Qk[j] * xr));
                    const T vbx = s.rbx[j] * zg[(cb + L::OBX + j) * ld] - yg[(cb + L::OBX + j) * ld];
                    T v = sigma * xg[(vb + L::OX + j) * ld] - qh - s.ex[j] * vd_cur[j] + s.bx[j] * vbx;
#pragma unroll
                    for (int i = 0; i < NX; ++i) v += s.Ah[i][j] * vd_next[i];
                    if (NS) {
                        const T rs = sigma * xg[(vb + L::OS + j) * ld] + s.bs[j] * vbx;
                        v -= (s.mxs[j] / s.mss[j]) * rs;
                    }
                    r[j] = v;
                }
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    T v = 0;
                    if (!last) {
                        const T vbu = s.rbu[j] * zg[(cb + L::OBU + j) * ld] - yg[(cb + L::OBU + j) * ld];
                        v = sigma * xg[(vb + L::OU + j) * ld] + s.bu[j] * vbu;
#pragma unroll
                        for (int i = 0; i < NX; ++i) v += s.Bh[i][j] * vd_next[i];
                    }
                    r[NX + j] = v;
                }
                if (k > 0) {
                    const size_t fp = (size_t)(k - 1) * L::FAC + L::LT;
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        T acc = 0;
#pragma unroll
                        for (int a = 0; a < NW; ++a) acc += fg[(fp + i * NW + a) * ld] * tprev[a];
                        r[i] -= acc;
                    }
                }
#pragma unroll
                for (int a = 0; a < NW; ++a) {
                    T acc = 0;
#pragma unroll
                    for (int d = 0; d <= a; ++d) acc += fg[(fb + a * (a + 1) / 2 + d) * ld] * r[d];
                    tprev[a] = acc;
                }
#pragma unroll
                for (int a = 0; a < NW; ++a) tg[((size_t)k * NW + a) * ld] = tprev[a];
#pragma unroll
                for (int i = 0; i < NX; ++i) { Ed_cur[i] = s.Ed_next[i]; vd_cur[i] = vd_next[i]; }
            }
        }
        // ================= backward sweep: L^{-T}, then x / z / y updates
        {
            T xt_next[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) xt_next[i] = 0;
            for (int k = N; k >= 0; --k) {
                const bool last = (k == N);
                const size_t vb = (size_t)k * L::VS, cb = (size_t)k * L::CS, fb = (size_t)k * L::FAC;
                if (p.tv && !last) load_model<T, L>(p, b, k, m);
                T Ed_cur[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) Ed_cur[i] = Eg[(cb + L::OD + i) * ld];
                stage_coef<T, L>(p, b, k, m, Ed_cur, c, rho, rho_eq, s);
                T rhs[NW], w[NW];
#pragma unroll
                for (int a = 0; a < NW; ++a) rhs[a] = tg[((size_t)k * NW + a) * ld];
                if (!last) {
#pragma unroll
                    for (int a = 0; a < NW; ++a) {
                        T acc = 0;
#pragma unroll
                        for (int i = 0; i < NX; ++i) acc += fg[(fb + L::LT + i * NW + a) * ld] * xt_next[i];
                        rhs[a] -= acc;
                    }
                }
#pragma unroll
                for (int d = 0; d < NW; ++d) {
                    T acc = 0;
#pragma unroll
                    for (int a = d; a < NW; ++a) acc += fg[(fb + a * (a + 1) / 2 + d) * ld] * rhs[a];
                    w[d] = acc;
                }
                // rows bx_k (+ slack recovery), x_k / s_k updates
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    T zb = zg[(cb + L::OBX + j) * ld], yb = yg[(cb + L::OBX + j) * ld];
                    T ztil = s.bx[j] * w[j];
                    if (NS) {
                        const T sold = xg[(vb + L::OS + j) * ld];
                        const T vbx = s.rbx[j] * zb - yb;
                        const T rs = sigma * sold + s.bs[j] * vbx;
                        const T st = (rs - s.mxs[j] * w[j]) / s.mss[j];
                        ztil += s.bs[j] * st;
                        xg[(vb + L::OS + j) * ld] = alpha * st + ((T)1 - alpha) * sold;
                    }
                    row_update(ztil, s.lbx[j], s.ubx[j], s.rbx[j], alpha, zb, yb);
                    zg[(cb + L::OBX + j) * ld] = zb;
                    yg[(cb + L::OBX + j) * ld] = yb;
                    const T xold = xg[(vb + L::OX + j) * ld];
                    xg[(vb + L::OX + j) * ld] = alpha * w[j] + ((T)1 - alpha) * xold;
                }
                if (!last) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        T zb = zg[(cb + L::OBU + j) * ld], yb = yg[(cb + L::OBU + j) * ld];
                        row_update(s.bu[j] * w[NX + j], s.lbu[j], s.ubu[j], s.rbu[j], alpha, zb, yb);
                        zg[(cb + L::OBU + j) * ld] = zb;
                        yg[(cb + L::OBU + j) * ld] = yb;
                        const T uold = xg[(vb + L::OU + j) * ld];
                        xg[(vb + L::OU + j) * ld] = alpha * w[NX + j] + ((T)1 - alpha) * uold;
                    }
                    // rows dyn_{k+1}:  Ah x~_k + Bh u~_k - ex_{k+1} x~_{k+1} = -E g_k
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T exn = s.Ed_next[i] * Dg[(vb + L::VS + L::OX + i) * ld];
                        T ztil = -exn * xt_next[i];
#pragma unroll
                        for (int j = 0; j < NX; ++j) ztil += s.Ah[i][j] * w[j];
#pragma unroll
                        for (int j = 0; j < NU; ++j) ztil += s.Bh[i][j] * w[NX + j];
                        const T beq = -s.Ed_next[i] * m.g[i];
                        T zb = zg[(cb + L::CS + L::OD + i) * ld], yb = yg[(cb + L::CS + L::OD + i) * ld];
                        row_update(ztil, beq, beq, rho_eq, alpha, zb, yb);
                        zg[(cb + L::CS + L::OD + i) * ld] = zb;
                        yg[(cb + L::CS + L::OD + i) * ld] = yb;
                    }
                }
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T beq = -Ed_cur[i] * xinit[i];
                        T zb = zg[(size_t)(L::OD + i) * ld], yb = yg[(size_t)(L::OD + i) * ld];
                        row_update(-s.ex[i] * w[i], beq, beq, rho_eq, alpha, zb, yb);
                        zg[(size_t)(L::OD + i) * ld] = zb;
                        yg[(size_t)(L::OD + i) * ld] = yb;
                    }
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) xt_next[i] = w[i];
            }
        }
        // ================= termination test on the unscaled residuals (auxil.c: check_termination)
        checked = 0;
        const bool at_check = p.check_every > 0 && (it % p.check_every == 0);
        if (at_check || it == p.max_iter) {
            checked = at_check;
            T nz = 0, nAx = 0, nq = 0, nAty = 0, nPx = 0;
            pri = 0; dua = 0;
            T Ed_cur[NX], yd_cur[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                Ed_cur[i] = Eg[(size_t)(L::OD + i) * ld];
                yd_cur[i] = yg[(size_t)(L::OD + i) * ld];
            }
            for (int k = 0; k <= N; ++k) {
                const bool last = (k == N);
                const size_t vb = (size_t)k * L::VS, cb = (size_t)k * L::CS;
                if (p.tv && !last) load_model<T, L>(p, b, k, m);
                stage_coef<T, L>(p, b, k, m, Ed_cur, c, rho, rho_eq, s);
                const T* Qk = last ? p.QN : p.Q;
                T xk[NX], uk[NU], yd_next[NX];
#pragma unroll
                for (int j = 0; j < NX; ++j) xk[j] = xg[(vb + L::OX + j) * ld];
#pragma unroll
                for (int j = 0; j < NU; ++j) uk[j] = last ? (T)0 : xg[(vb + L::OU + j) * ld];
#pragma unroll
                for (int i = 0; i < NX; ++i) yd_next[i] = last ? (T)0 : yg[(cb + L::CS + L::OD + i) * ld];
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T Einv = (T)1 / Ed_cur[i];
                        const T ax = -s.ex[i] * xk[i], zz = zg[(size_t)(L::OD + i) * ld];
                        pri = tmax(pri, tabs(Einv * (ax - zz)));
                        nz = tmax(nz, tabs(Einv * zz));
                        nAx = tmax(nAx, tabs(Einv * ax));
                    }
                }
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    const T Dj = Dg[(vb + L::OX + j) * ld], Dinv = (T)1 / Dj;
                    const T ybx = yg[(cb + L::OBX + j) * ld], zbx = zg[(cb + L::OBX + j) * ld];
                    const T Ebx = Eg[(cb + L::OBX + j) * ld], Einv = (T)1 / Ebx;
                    T sk = 0;
                    if (NS) sk = xg[(vb + L::OS + j) * ld];
                    const T ax = s.bx[j] * xk[j] + s.bs[j] * sk;
                    pri = tmax(pri, tabs(Einv * (ax - zbx)));
                    nz = tmax(nz, tabs(Einv * zbx));
                    nAx = tmax(nAx, tabs(Einv * ax));
                    const T xr = p.Xr[((p.xr_tv ? (size_t)k * NX : 0) + j) * ld + b];
                    const T qh = c * Dj * (-(Qk[j] * xr));
                    T aty = -s.ex[j] * yd_cur[j] + s.bx[j] * ybx;
#pragma unroll
                    for (int i = 0; i < NX; ++i) aty += s.Ah[i][j] * yd_next[i];
                    const T px = s.px[j] * xk[j];
                    dua = tmax(dua, tabs(Dinv * (qh + aty + px)));
                    nq = tmax(nq, tabs(Dinv * qh));
                    nAty = tmax(nAty, tabs(Dinv * aty));
                    nPx = tmax(nPx, tabs(Dinv * px));
                    if (NS) {
                        const T Dsinv = (T)1 / Dg[(vb + L::OS + j) * ld];
                        const T atys = s.bs[j] * ybx, pxs = s.ps[j] * sk;
                        dua = tmax(dua, tabs(Dsinv * (atys + pxs)));
                        nAty = tmax(nAty, tabs(Dsinv * atys));
                        nPx = tmax(nPx, tabs(Dsinv * pxs));
                    }
                }
                if (!last) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        const T Dinv = (T)1 / Dg[(vb + L::OU + j) * ld];
                        const T ybu = yg[(cb + L::OBU + j) * ld], zbu = zg[(cb + L::OBU + j) * ld];
                        const T Einv = (T)1 / Eg[(cb + L::OBU + j) * ld];
                        const T ax = s.bu[j] * uk[j];
                        pri = tmax(pri, tabs(Einv * (ax - zbu)));
                        nz = tmax(nz, tabs(Einv * zbu));
                        nAx = tmax(nAx, tabs(Einv * ax));
                        T aty = s.bu[j] * ybu;
#pragma unroll
                        for (int i = 0; i < NX; ++i) aty += s.Bh[i][j] * yd_next[i];
                        const T px = s.pu[j] * uk[j];
                        dua = tmax(dua, tabs(Dinv * (aty + px)));
                        nAty = tmax(nAty, tabs(Dinv * aty));
                        nPx = tmax(nPx, tabs(Dinv * px));
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T Einv = (T)1 / s.Ed_next[i];
                        const T exn = s.Ed_next[i] * Dg[(vb + L::VS + L::OX + i) * ld];
                        T ax = -exn * xg[(vb + L::VS + L::OX + i) * ld];
#pragma unroll
                        for (int j = 0; j < NX; ++j) ax += s.Ah[i][j] * xk[j];
#pragma unroll
                        for (int j = 0; j < NU; ++j) ax += s.Bh[i][j] * uk[j];
                        const T zz = zg[(cb + L::CS + L::OD + i) * ld];
                        pri = tmax(pri, tabs(Einv * (ax - zz)));
                        nz = tmax(nz, tabs(Einv * zz));
                        nAx = tmax(nAx, tabs(Einv * ax));
                    }
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) { Ed_cur[i] = s.Ed_next[i]; yd_cur[i] = yd_next[i]; }
            }
            dua *= cinv;
            const T mp = tmax(nz, nAx);
            const T md = cinv * tmax(nq, tmax(nAty, nPx));
            if (at_check) {
                if (pri < p.eps_abs + p.eps_rel * mp && dua < p.eps_abs + p.eps_rel * md) {
                    status = kSolved;
                    break;
                }
            }
            if (it == p.max_iter) {
                // osqp.c (end of osqp_solve): exact test if not yet done this iteration, then the
                // approximate test (tolerances x10) before declaring max-iter
                if (pri < p.eps_abs + p.eps_rel * mp && dua < p.eps_abs + p.eps_rel * md)
                    status = kSolved;
                else if (pri < (T)10 * (p.eps_abs + p.eps_rel * mp) && dua < (T)10 * (p.eps_abs + p.eps_rel * md))
                    status = kSolvedInaccurate;
                else
                    status = kMaxIterReached;
                break;
            }
        }
    }
    (void)checked;
    p.iter[b] = it > p.max_iter ? p.max_iter : it;
    p.status[b] = status;
    p.pri_res[b] = pri;
    p.dua_res[b] = dua;
}

}  // namespace mpcb
