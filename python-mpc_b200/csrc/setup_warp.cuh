// Ruiz equilibration with a warp per QP: lane k owns stage k.
//
// A Ruiz pass is a Jacobi step — every new scaling depends on the OLD scalings of its own stage and of the two neighbour
// stages only (scale_stage_update, qp_thread.cuh) — so the stages of a pass are independent.  The lane-per-QP kernel
// (scale_one) walks them one after the other and ping-pongs D, E through HBM ten times (4.8 GB for 65536 QPs of the
// configs[2] shape); here the scalings of a stage stay in its lane's REGISTERS over all passes, the neighbour values come
// by warp shuffle, the two statistics of the cost normalisation are reduced over the lanes, and D, E, c are written ONCE
// (0.24 GB).  Horizons up to 31 stages (N + 1 <= 32); longer ones keep the lane-per-QP kernel.
#pragma once
#include "qp_thread.cuh"

#if defined(__CUDACC__) && !defined(MPCB_EMU)
namespace mpcb {

constexpr int SCALE_WARPS = 8;       // QPs per CTA

template <typename T, typename L>
__global__ void __launch_bounds__(SCALE_WARPS * 32) scale_warp_kernel(const __grid_constant__ KParams<T> p) {
    constexpr int NX = L::NX, NU = L::NU, NS = L::NS;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int b = blockIdx.x * SCALE_WARPS + warp;
    if (b >= p.B) return;
    const int N = p.N;
    const bool stage = k <= N;
    const int kk = stage ? k : N;                   // lanes beyond the horizon shadow the last stage, never store
    Model<T, L> m;
    load_model<T, L>(p, b, p.tv ? (kk < N ? kk : N - 1) : 0, m);
    ScaleStage<T, L> o, n;
    T E0[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { o.Dx[i] = o.Dsl[i] = o.Ebx[i] = o.Edn[i] = (T)1; E0[i] = (T)1; }
#pragma unroll
    for (int j = 0; j < NU; ++j) { o.Du[j] = o.Ebu[j] = (T)1; }
    T c = (T)1;
    for (int it = 0; it < p.scaling; ++it) {
        T Ed_cur[NX], Dx_next[NX], E0new[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const T up = __shfl_up_sync(0xffffffffu, o.Edn[i], 1), dn = __shfl_down_sync(0xffffffffu, o.Dx[i], 1);
            Ed_cur[i] = k == 0 ? E0[i] : up;        // E of rows dyn_k: the previous stage's (dyn_0: the header's)
            Dx_next[i] = k >= N ? (T)1 : dn;
            E0new[i] = E0[i];
        }
        T sumP = 0, maxq = 0;
        if (stage) scale_stage_update<T, L>(p, m, c, b, k, o, Ed_cur, Dx_next, n, E0new, sumP, maxq);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sumP += __shfl_xor_sync(0xffffffffu, sumP, off);
            maxq = tmax(maxq, __shfl_xor_sync(0xffffffffu, maxq, off));
        }
        c = scale_cost_update<T, L>(c, sumP, maxq, N);
        if (stage) o = n;
#pragma unroll
        for (int i = 0; i < NX; ++i) E0[i] = E0new[i];
    }
    if (!stage) return;
    Ws<T, L> ws(p, b);
    T* R = ws.R(k);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        MPCB_AT(R, L::R_D + L::OX + i) = o.Dx[i];
        if (NS) MPCB_AT(R, L::R_D + L::OS + (NS ? i : 0)) = o.Dsl[i];
        MPCB_AT(R, L::R_E + L::ODN + i) = o.Edn[i];
        MPCB_AT(R, L::R_E + L::OBX + i) = o.Ebx[i];
        if (k == 0) MPCB_AT(ws.hdr, L::H_E0 + i) = E0[i];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        MPCB_AT(R, L::R_D + L::OU + j) = o.Du[j];
        MPCB_AT(R, L::R_E + L::OBU + j) = o.Ebu[j];
    }
    if (k == 0) MPCB_AT(ws.hdr, L::H_C) = c;
}

}  // namespace mpcb
#endif
