// Shared declarations for the mpc_b200 kernels (device + host).
//
// Vocabulary (follows the reference, Control/MPC/*.py): a *QP* is one MPC problem over a
// horizon of N *stages*; its variables are the stage states x_k, inputs u_k and (soft
// constraint formulation) slacks s_k; its rows are the dynamics rows dyn_k and the bound
// rows bx_k / bu_k.  A *batch* is B independent QPs.
//
// Data layout in HBM.  Every per-QP array is stored "element-major" (SoA):
//     a[e * ld + b]      e = element index inside one QP,  b = QP index,  ld >= B
// so that a warp whose lanes own 32 consecutive QPs reads 128 contiguous bytes per element.
// Inside one QP, vectors are stored *stage-major* (all of stage 0, then stage 1 ...):
//     variables   stage k at k*VS :  [ x_k (NX) | s_k (NS) | u_k (NU) ]       VS = NX+NS+NU
//     rows        stage k at k*CS :  [ dyn_k (NX) | bx_k (NX) | bu_k (NU) ]   CS = 2NX+NU
// (the u / bu slots of stage N are unused padding).  The reference's ordering
// (x_0..x_N, u_0..u_{N-1}, s_0..s_N) is produced by the gather kernels in mpc_b200.cu.
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define MPCB_HD __host__ __device__ __forceinline__
#else
#define MPCB_HD inline
#endif

namespace mpcb {

constexpr int MAXNX = 10;
constexpr int MAXNU = 2;

// osqp/include/constants.h (0.6.x) — the same constants as oracle/osqp_admm.{py,c}
constexpr double kOsqpInfty = 1e30;
constexpr double kRhoMin = 1e-6;
constexpr double kRhoMax = 1e6;
constexpr double kRhoEqOverRhoIneq = 1e3;
constexpr double kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4;
constexpr double kMaxScaling = 1e4;

enum Status : int {
    kSolved = 1,
    kSolvedInaccurate = 2,
    kPrimalInfeasibleInaccurate = 3,
    kDualInfeasibleInaccurate = 4,
    kMaxIterReached = -2,
    kPrimalInfeasible = -3,
    kDualInfeasible = -4,
    kUnsolved = -10,
};

// Everything a per-QP kernel needs, passed by value as the kernel argument.
template <typename T>
struct KParams {
    int N;            // horizon
    int B;            // QPs in this launch
    size_t ld;        // leading dimension of every SoA array
    // ---- model (inputs): A_k (NX*NX), B_k (NX*NU), g_k (NX), row-major inside a stage
    const T* Ad;
    const T* Bd;
    const T* gd;      // may be null (zero)
    int tv;           // 1: one (A,B,g) per stage, stage stride = element count; 0: one per QP
    int model_bs;     // batch stride of the model arrays: 1 per-QP linearisation, 0 one shared linearisation
    const T* x_init;  // (NX)
    const T* Xr;      // (NX) or (N+1)*(NX)
    int xr_tv;
    // ---- weights and bounds, shared by the whole batch (constructor arguments of the controller)
    T Q[MAXNX], QN[MAXNX], R[MAXNU], W[MAXNX], S[MAXNX];
    T xmin[MAXNX], xmax[MAXNX], umin[MAXNU], umax[MAXNU];
    const T* xbox;    // optional per-stage state bounds, shared: [(N+1)][2][NX] (mpc_ of mpc_kinematics.py:215); null -> xmin/xmax
    // ---- OSQP settings
    T rho, sigma, alpha, eps_abs, eps_rel, eps_pinf, eps_dinf;
    int max_iter, scaling, check_every;
    int warm;         // 0: cold start (x = z = y = 0); 1: keep the iterates already in the workspace
    // ---- workspace (per QP, SoA)
    T* D;             // [2][(N+1)*VS]   ping-pong during Ruiz, result in half 0
    T* E;             // [2][(N+1)*CS]
    T* c;             // [1]
    T* fac;           // [(N+1)*FAC]     block-bidiagonal Cholesky factor (inverse diagonal blocks + coupling blocks)
    T* x;             // [(N+1)*VS]      scaled iterates
    T* z;             // [(N+1)*CS]
    T* y;             // [(N+1)*CS]
    T* t;             // [(N+1)*NW]      forward-substitution intermediate
    int* iter;        // [B]
    int* status;      // [B]
    T* pri_res;       // [B]
    T* dua_res;       // [B]
};

template <int NX_, int NU_, bool SLACK_>
struct Lay {
    static constexpr int NX = NX_, NU = NU_;
    static constexpr bool SLACK = SLACK_;
    static constexpr int NS = SLACK_ ? NX_ : 0;
    static constexpr int NW = NX_ + NU_;
    static constexpr int VS = NX_ + NS + NU_;
    static constexpr int CS = 2 * NX_ + NU_;
    static constexpr int LT = NW * (NW + 1) / 2;
    static constexpr int FS = NX_ * NW;
    static constexpr int FAC = LT + FS;
    static constexpr int OX = 0, OS = NX_, OU = NX_ + NS;     // variable offsets inside a stage
    static constexpr int OD = 0, OBX = NX_, OBU = 2 * NX_;    // row offsets inside a stage
    static MPCB_HD int nvar(int N) { return (N + 1) * NX + N * NU + (N + 1) * NS; }
    static MPCB_HD int ncon(int N) { return 2 * (N + 1) * NX + N * NU; }
};

}  // namespace mpcb
