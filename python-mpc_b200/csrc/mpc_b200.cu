// mpc_b200 — kernels, launch layer and C ABI (include/mpc_b200.h).
//
// Compiled by nvcc for sm_100a into python-mpc_b200/libmpc_b200.so (the product).
// tests/emu compiles this same file with g++ and -DMPCB_EMU: kernel launches become host loops
// and the CUDA runtime calls become malloc/memcpy, so that the whole C ABI can be exercised in
// the GPU-less build container.  That build is test infrastructure only; the product loader
// (python-mpc_b200/_lib.py) never looks for it.
#include "../../include/mpc_b200.h"
#include "mpc_common.h"
#include "models.cuh"
#include "qp_thread.cuh"
#include "admm_kernel.cuh"
#include "admm_wide.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>

using namespace mpcb;

// =============================================================================================
// runtime layer
// =============================================================================================
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int env_int(const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; }
static std::atomic<int> g_opt_tma{std::getenv("MPCB_NO_TMA") ? 0 : 1};
static std::atomic<int> g_opt_retile{std::getenv("MPCB_NO_RETILE") ? 0 : 1};
static std::atomic<int> g_opt_cert{std::getenv("MPCB_NO_CERT") ? 0 : 1};
static std::atomic<int> g_opt_wide{std::getenv("MPCB_NO_WIDE") ? 0 : 1};
static std::atomic<int> g_opt_retile_min{env_int("MPCB_RETILE_MIN_BATCH", 4096)};

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#ifndef MPCB_EMU
#include <cuda_runtime.h>
typedef cudaStream_t rt_stream;
#define RT_CHECK(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(MPCB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
    } while (0)
static int rt_malloc(void** p, size_t n) { RT_CHECK(cudaMalloc(p, n ? n : 1)); return 0; }
static void rt_free(void* p) { if (p) cudaFree(p); }
static int rt_memset(void* p, int v, size_t n, rt_stream s) { RT_CHECK(cudaMemsetAsync(p, v, n, s)); return 0; }
static int rt_h2d(void* d, const void* h, size_t n, rt_stream s) { RT_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); return 0; }
static int rt_d2h(void* h, const void* d, size_t n, rt_stream s) { RT_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); return 0; }
static int rt_sync(rt_stream s) { RT_CHECK(cudaStreamSynchronize(s)); return 0; }
// MPCB_DEBUG_SYNC=1: synchronise after every launch so that a faulting kernel is named in the error
static bool debug_sync() {
    static const bool on = std::getenv("MPCB_DEBUG_SYNC") != nullptr;
    return on;
}
static int rt_launch_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug_sync()) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(MPCB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}

#ifndef MPCB_QP_THREADS
#define MPCB_QP_THREADS 128     // threads per CTA of the per-QP kernels
#endif
#ifndef MPCB_QP_MINBLOCKS
#define MPCB_QP_MINBLOCKS 2     // resident CTAs per SM the register allocation must allow
#endif
template <typename Op, typename T, typename L>
__global__ void __launch_bounds__(MPCB_QP_THREADS, MPCB_QP_MINBLOCKS) qp_kernel(const KParams<T> p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < p.B) Op::template run<T, L>(p, b);
}
template <typename Op, typename T, typename L>
static int launch_qp(const KParams<T>& p, rt_stream st) {
    const int threads = MPCB_QP_THREADS;
    qp_kernel<Op, T, L><<<(p.B + threads - 1) / threads, threads, 0, st>>>(p);
    ++g_launches;
    return rt_launch_check(Op::name());
}
template <typename F>
__global__ void __launch_bounds__(128) lambda_kernel(int n, F f) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) f(b);
}
#define MPCB_LAMBDA [=] __device__
template <typename F>
static int launch_1d(int n, rt_stream st, F f) {
    if (n <= 0) return 0;
    lambda_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, f);
    ++g_launches;
    return rt_launch_check("lambda_kernel");
}
#else   // ---------------------------------------------------------------- MPCB_EMU (tests only)
typedef void* rt_stream;
static int rt_malloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : fail(MPCB_E_ALLOC, "malloc"); }
static void rt_free(void* p) { std::free(p); }
static int rt_memset(void* p, int v, size_t n, rt_stream) { std::memset(p, v, n); return 0; }
static int rt_h2d(void* d, const void* h, size_t n, rt_stream) { std::memcpy(d, h, n); return 0; }
static int rt_d2h(void* h, const void* d, size_t n, rt_stream) { std::memcpy(h, d, n); return 0; }
static int rt_sync(rt_stream) { return 0; }
template <typename Op, typename T, typename L>
static int launch_qp(const KParams<T>& p, rt_stream) {
    for (int b = 0; b < p.B; ++b) Op::template run<T, L>(p, b);
    ++g_launches;
    return 0;
}
#define MPCB_LAMBDA [=]
template <typename F>
static int launch_1d(int n, rt_stream, F f) {
    for (int b = 0; b < n; ++b) f(b);
    ++g_launches;
    return 0;
}
#endif

// =============================================================================================
// per-QP kernel bodies
// =============================================================================================
struct ScaleOp { static const char* name() { return "scale"; } template <typename T, typename L> static MPCB_HD void run(const KParams<T>& p, int b) { scale_one<T, L>(p, b); } };
struct FactorOp { static const char* name() { return "factor"; } template <typename T, typename L> static MPCB_HD void run(const KParams<T>& p, int b) { factor_one<T, L>(p, b); } };
struct AdmmOp { static const char* name() { return "admm"; } template <typename T, typename L> static MPCB_HD void run(const KParams<T>& p, int b) { admm_one<T, L>(p, b); } };
struct ColdOp { static const char* name() { return "cold_start"; } template <typename T, typename L> static MPCB_HD void run(const KParams<T>& p, int b) { Ws<T, L> ws(p, b); admm_cold_start<T, L>(p, ws); } };

// Explicit QP data in the reference's ordering (see mpcb_build_qp in the header).
struct BuildOut {
    void *Pdiag, *q, *Avals, *l, *u;
};
template <typename T, typename L>
MPCB_HD void build_one(const KParams<T>& p, const BuildOut& o, int b) {
    constexpr int NX = L::NX, NU = L::NU, NS = L::NS;
    const int N = p.N;
    const size_t ld = p.ld;
    T* Pd = (T*)o.Pdiag; T* q = (T*)o.q; T* Av = (T*)o.Avals; T* lo = (T*)o.l; T* up = (T*)o.u;
    const size_t ux0 = (size_t)(N + 1) * NX, sx0 = ux0 + (size_t)N * NU;      // variable offsets
    const size_t bx0 = (size_t)(N + 1) * NX, bu0 = 2 * (size_t)(N + 1) * NX;  // row offsets
    // CSC value offsets: x columns of stage k<N hold (2+NX) values, of stage N hold 2; u columns NX+1; s columns 1
    const size_t nnz_x = (size_t)N * NX * (2 + NX) + (size_t)NX * 2;
    const size_t nnz_u = (size_t)N * NU * (NX + 1);
    Model<T, L> m;
    if (!p.tv) load_model<T, L>(p, b, 0, m);
    for (int k = 0; k <= N; ++k) {
        const bool last = (k == N);
        if (p.tv && !last) load_model<T, L>(p, b, k, m);
        const T* Qk = last ? p.QN : p.Q;
        T blo[NX], bhi[NX];
        for (int i = 0; i < NX; ++i) {
            blo[i] = p.xbox ? p.xbox[(k * 2 + 0) * NX + i] : p.xmin[i];
            bhi[i] = p.xbox ? p.xbox[(k * 2 + 1) * NX + i] : p.xmax[i];
        }
        for (int j = 0; j < NX; ++j) {
            const size_t v = (size_t)k * NX + j;
            const T xr = p.Xr[((p.xr_tv ? (size_t)k * NX : 0) + j) * ld + b];
            if (Pd) Pd[v * ld + b] = Qk[j];
            if (q) q[v * ld + b] = -(Qk[j] * xr);
            if (Av) {
                size_t a = (size_t)k * NX * (2 + NX) + (size_t)j * (last ? 2 : 2 + NX);
                Av[a++ * ld + b] = (T)-1;
                if (!last)
                    for (int i = 0; i < NX; ++i) Av[a++ * ld + b] = m.A[i][j];
                Av[a * ld + b] = (T)1;
            }
            // rows dyn_k, bx_k
            const T beq = k == 0 ? -p.x_init[(size_t)j * ld + b] : (T)0;   // dyn_k for k>0 written below via g_{k-1}
            if (k == 0) { if (lo) lo[v * ld + b] = beq; if (up) up[v * ld + b] = beq; }
            if (!last) {
                const size_t r = (size_t)(k + 1) * NX + j;
                if (lo) lo[r * ld + b] = -model_g<T, L>(p, b, k, j);
                if (up) up[r * ld + b] = -model_g<T, L>(p, b, k, j);
            }
            if (lo) lo[(bx0 + v) * ld + b] = blo[j];
            if (up) up[(bx0 + v) * ld + b] = bhi[j];
            if (NS) {
                const size_t sv = sx0 + v;
                if (Pd) Pd[sv * ld + b] = p.W[j];
                if (q) q[sv * ld + b] = (T)0;
                if (Av) Av[(nnz_x + nnz_u + v) * ld + b] = p.S[j];
            }
        }
        if (!last) {
            for (int j = 0; j < NU; ++j) {
                const size_t v = ux0 + (size_t)k * NU + j;
                if (Pd) Pd[v * ld + b] = p.R[j];
                if (q) q[v * ld + b] = (T)0;
                if (Av) {
                    size_t a = nnz_x + ((size_t)k * NU + j) * (NX + 1);
                    for (int i = 0; i < NX; ++i) Av[a++ * ld + b] = m.B[i][j];
                    Av[a * ld + b] = (T)1;
                }
                if (lo) lo[(bu0 + (size_t)k * NU + j) * ld + b] = p.umin[j];
                if (up) up[(bu0 + (size_t)k * NU + j) * ld + b] = p.umax[j];
            }
        }
    }
}

template <typename T, typename L>
struct BuildFn {
    KParams<T> p;
    BuildOut o;
    MPCB_HD void operator()(int b) const { build_one<T, L>(p, o, b); }
};

// =============================================================================================
// solver object
// =============================================================================================
struct mpcb_solver {
    mpcb_problem prob;
    mpcb_settings set;
    int cap = 0, batch = 0;
    size_t ld = 0;
    bool is_setup = false;
    size_t esz = 4;
    // workspace
    void *rec = nullptr, *hdr = nullptr, *yrows = nullptr, *scr = nullptr, *scr_hdr = nullptr;
    void *pri = nullptr, *dua = nullptr, *xbox = nullptr, *xbox_alt = nullptr;   // xbox_alt: the other buffer of an update
    int inf_prob = 0, inf_box = 0;      // an infinite bound among the constructor bounds / the per-stage boxes
    int *iter = nullptr, *status = nullptr, *tile_counter = nullptr;
    // re-tiling of unconverged QPs (see run_admm): survivor lists and a half-size scratch workspace
    int *surv[2] = {nullptr, nullptr}, *n_surv = nullptr, *tile_prog = nullptr;
    int retile_at = 0;          // iteration count at which the previous solve re-tiled (0: not known yet)
    void *rec2 = nullptr, *hdr2 = nullptr, *yrows2 = nullptr;
    size_t ld2 = 0;
    // borrowed inputs
    const void *Ad = nullptr, *Bd = nullptr, *gd = nullptr, *x_init = nullptr, *Xr = nullptr;
    // staging for the host front door
    void* stage_in = nullptr; size_t stage_in_bytes = 0;
    void* stage_out = nullptr; size_t stage_out_bytes = 0;
    void* soa_in = nullptr; size_t soa_in_bytes = 0;
    size_t ws_bytes = 0;
    int VS = 0, CS = 0, NW = 0, LT = 0, REC = 0, HDR = 0, nvar = 0, ncon = 0;
    int inf_bounds = 0;
    // prob.setup() leaves x = z = y = 0 (osqp.c: osqp_setup -> cold_start): the first ADMM launch after a setup starts
    // cold whatever warm_start says; readers of the iterates before that launch get the zeros written on demand
    bool cold_pending = false;
    int dev = 0, dev_max_smem = 0, dev_sms = 0;      // device the workspace lives on and its launch-sizing attributes
};

template <typename T>
static KParams<T> make_params(const mpcb_solver* s) {
    KParams<T> p;
    std::memset(&p, 0, sizeof(p));
    const mpcb_problem& q = s->prob;
    p.N = q.horizon; p.B = s->batch; p.ld = s->ld;
    p.Ad = (const T*)s->Ad; p.Bd = (const T*)s->Bd; p.gd = (const T*)s->gd;
    p.tv = q.time_varying; p.model_bs = q.shared_model ? 0 : 1;
    p.x_init = (const T*)s->x_init; p.Xr = (const T*)s->Xr; p.xr_tv = q.stage_reference;
    for (int i = 0; i < MAXNX; ++i) {
        p.Q[i] = (T)q.Q[i]; p.QN[i] = (T)q.QN[i]; p.W[i] = (T)q.W[i]; p.S[i] = (T)q.S[i];
        p.xmin[i] = (T)clip_infty(q.xmin[i]); p.xmax[i] = (T)clip_infty(q.xmax[i]);     // python interface of OSQP: +-inf -> +-OSQP_INFTY
    }
    for (int i = 0; i < MAXNU; ++i) {
        p.R[i] = (T)q.R[i]; p.umin[i] = (T)clip_infty(q.umin[i]); p.umax[i] = (T)clip_infty(q.umax[i]);
    }
    p.xbox = (const T*)s->xbox;
    p.inf_bounds = s->inf_bounds;
    p.certs = g_opt_cert.load();
    const mpcb_settings& o = s->set;
    p.rho = (T)o.rho; p.sigma = (T)o.sigma; p.alpha = (T)o.alpha; p.eps_abs = (T)o.eps_abs; p.eps_rel = (T)o.eps_rel;
    p.eps_pinf = (T)o.eps_prim_inf; p.eps_dinf = (T)o.eps_dual_inf;
    p.max_iter = o.max_iter; p.scaling = o.scaling; p.check_every = o.check_termination; p.warm = o.warm_start;
    p.rec = (T*)s->rec; p.hdr = (T*)s->hdr; p.yrows = (T*)s->yrows; p.scr = (T*)s->scr; p.scr_hdr = (T*)s->scr_hdr;
    p.iter = s->iter; p.status = s->status; p.pri_res = (T*)s->pri; p.dua_res = (T*)s->dua;
    p.it0 = 0; p.it_stop = o.max_iter; p.qp_map = nullptr; p.survivors = s->surv[0]; p.n_survivors = s->n_surv;
    p.chunk_len = o.check_termination; p.tile_prog = s->tile_prog; p.list_survivors = 0;
    return p;
}

// ---- shape dispatch: the (nx, nu, slack) combinations of the reference's formulations
//   lateral (4,1): vanilla / slack;  lateral delta-u (5,1): plain / slack  (vehicle_lateral_mpc_slack_increment.py)
//   kinematic (4,2) and its delta-u form (6,2)   (mpc_kinematics*.py, mpc_incre_kine_func.py)
//   dynamic (6,2) and its delta-u form (8,2)     (mpc_dynamics.py)
#ifdef MPCB_DEV_SHAPE    // development build (scripts/devbuild.sh): only the configs[2] shape, compiles in seconds
#define MPCB_SHAPES(X) X(5, 1, true)
#else
#define MPCB_SHAPES(X) X(4, 1, false) X(4, 1, true) X(5, 1, false) X(5, 1, true) X(4, 2, false) X(6, 2, false) X(8, 2, false)
#endif

template <typename Fn>
static int dispatch(const mpcb_solver* s, Fn&& fn) {
    const mpcb_problem& q = s->prob;
#define X(NX_, NU_, SL_)                                                              \
    if (q.nx == NX_ && q.nu == NU_ && (q.slack != 0) == SL_) {                        \
        typedef Lay<NX_, NU_, SL_> L;                                                 \
        if (q.dtype == MPCB_F32) return fn((float*)nullptr, (L*)nullptr);             \
        return fn((double*)nullptr, (L*)nullptr);                                     \
    }
    MPCB_SHAPES(X)
#undef X
    return fail(MPCB_E_ARG, "unsupported (nx, nu, slack) combination");
}

// anything this large can scale past the OSQP_INFTY test of set_rho_vec (E is at least MIN_SCALING)
static bool is_big(double v) { return !(std::fabs(v) < kOsqpInfty * kMinScaling * 1e-6); }
static int problem_has_inf_bounds(const mpcb_problem& q) {
    for (int i = 0; i < q.nx; ++i) if (is_big(q.xmin[i]) || is_big(q.xmax[i])) return 1;
    for (int i = 0; i < q.nu; ++i) if (is_big(q.umin[i]) || is_big(q.umax[i])) return 1;
    return 0;
}

static bool shape_supported(int nx, int nu, int slack) {
#define X(NX_, NU_, SL_) if (nx == NX_ && nu == NU_ && (slack != 0) == SL_) return true;
    MPCB_SHAPES(X)
#undef X
    return false;
}

// All mpcb_* functions below are declared extern "C" by include/mpc_b200.h.

const char* mpcb_last_error(void) { return g_err.c_str(); }
int mpcb_version(void) { return 100; }
long long mpcb_launch_count(void) { return g_launches.load(); }

int mpcb_set_option(const char* name, int value) {
    if (!name) return fail(MPCB_E_ARG, "null option name");
    const std::string n(name);
    if (n == "tma") g_opt_tma = value != 0;
    else if (n == "retile") g_opt_retile = value != 0;
    else if (n == "certificates") g_opt_cert = value != 0;
    else if (n == "wide") g_opt_wide = value != 0;
    else if (n == "retile_min_batch") g_opt_retile_min = value;
    else return fail(MPCB_E_ARG, "unknown option: " + n);
    return 0;
}

void mpcb_default_settings(mpcb_settings* s) {
    s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
    s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->max_iter = 4000; s->scaling = 10;
    s->check_termination = 25; s->warm_start = 1;
}

static int check_settings(const mpcb_settings* o) {
    if (!(o->rho > 0) || !(o->sigma > 0) || !(o->alpha > 0 && o->alpha < 2) || o->max_iter <= 0 ||
        o->scaling < 0 || o->check_termination < 0 || o->eps_abs < 0 || o->eps_rel < 0 ||
        (o->eps_abs == 0 && o->eps_rel == 0))
        return fail(MPCB_E_ARG, "invalid settings (rho, sigma > 0; 0 < alpha < 2; max_iter > 0; eps >= 0, not both 0)");
    return 0;
}

int mpcb_create(const mpcb_problem* prob, const mpcb_settings* settings, int capacity, mpcb_solver** out) {
    if (!prob || !out || capacity <= 0) return fail(MPCB_E_ARG, "null problem / non-positive capacity");
    if (prob->horizon < 1) return fail(MPCB_E_ARG, "horizon must be >= 1");
    if (prob->dtype != MPCB_F32 && prob->dtype != MPCB_F64) return fail(MPCB_E_ARG, "dtype must be MPCB_F32 or MPCB_F64");
    if (!shape_supported(prob->nx, prob->nu, prob->slack))
        return fail(MPCB_E_ARG, "unsupported (nx, nu, slack) combination");
    for (int i = 0; i < prob->nx; ++i)
        if (prob->xmin[i] > prob->xmax[i]) return fail(MPCB_E_ARG, "lower bound must be lower than or equal to upper bound");
    for (int i = 0; i < prob->nu; ++i)
        if (prob->umin[i] > prob->umax[i]) return fail(MPCB_E_ARG, "lower bound must be lower than or equal to upper bound");
    mpcb_settings def;
    mpcb_default_settings(&def);
    if (!settings) settings = &def;
    if (int rc = check_settings(settings)) return rc;
    mpcb_solver* s = new (std::nothrow) mpcb_solver();
    if (!s) return fail(MPCB_E_ALLOC, "out of host memory");
    s->prob = *prob; s->set = *settings; s->cap = capacity;
#ifndef MPCB_EMU
    if (cudaGetDevice(&s->dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&s->dev_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&s->dev_sms, cudaDevAttrMultiProcessorCount, s->dev) != cudaSuccess) {
        const std::string msg = std::string("cuda device query: ") + cudaGetErrorString(cudaGetLastError());
        delete s;
        return fail(MPCB_E_CUDA, msg);
    }
#endif
    s->inf_prob = problem_has_inf_bounds(*prob);
    s->inf_bounds = s->inf_prob;
    s->esz = prob->dtype == MPCB_F32 ? 4 : 8;
    const int N = prob->horizon, nx = prob->nx, nu = prob->nu, ns = prob->slack ? nx : 0;
    s->VS = nx + ns + nu; s->CS = 2 * nx + nu; s->NW = nx + nu; s->LT = s->NW * (s->NW + 1) / 2;
    s->REC = 2 * s->VS + 2 * s->CS + s->LT + s->NW; s->HDR = 3 * nx + 1;
    s->nvar = (N + 1) * nx + N * nu + (N + 1) * ns; s->ncon = 2 * (N + 1) * nx + N * nu;
    s->ld = ((size_t)capacity + 31) / 32 * 32;
    const size_t ld = s->ld, e = s->esz, S1 = (size_t)(N + 1);
    struct { void** p; size_t n; } allocs[] = {
        {&s->rec, S1 * s->REC * ld * e}, {&s->hdr, (size_t)s->HDR * ld * e}, {&s->yrows, S1 * s->CS * ld * e},
        {&s->scr, S1 * (s->VS + s->CS) * ld * e}, {&s->scr_hdr, (size_t)nx * ld * e}, {&s->pri, ld * e},
        {&s->dua, ld * e}, {(void**)&s->iter, ld * sizeof(int)}, {(void**)&s->status, ld * sizeof(int)},
        {(void**)&s->tile_counter, 64}, {(void**)&s->surv[0], ld * sizeof(int)}, {(void**)&s->surv[1], ld * sizeof(int)},
        {(void**)&s->n_surv, 64}, {(void**)&s->tile_prog, (ld / 32 + 1) * sizeof(int)}};
    for (auto& a : allocs) {
        if (int rc = rt_malloc(a.p, a.n)) { mpcb_destroy(s); return rc; }
        s->ws_bytes += a.n;
    }
    rt_memset(s->rec, 0, S1 * s->REC * ld * e, 0);
    rt_memset(s->hdr, 0, (size_t)s->HDR * ld * e, 0);
    rt_memset(s->yrows, 0, S1 * s->CS * ld * e, 0);
    rt_memset(s->status, 0, ld * sizeof(int), 0);
    rt_sync(0);
    *out = s;
    return 0;
}

void mpcb_destroy(mpcb_solver* s) {
    if (!s) return;
    void* ptrs[] = {s->rec, s->hdr, s->yrows, s->scr, s->scr_hdr, s->pri, s->dua, s->iter, s->status, s->tile_counter,
                    s->surv[0], s->surv[1], s->n_surv, s->tile_prog, s->rec2, s->hdr2, s->yrows2,
                    s->xbox, s->xbox_alt, s->stage_in, s->stage_out, s->soa_in};
    for (void* p : ptrs) rt_free(p);
    delete s;
}

int mpcb_set_settings(mpcb_solver* s, const mpcb_settings* o) {
    if (!s || !o) return fail(MPCB_E_ARG, "null argument");
    if (int rc = check_settings(o)) return rc;
    const bool refactor = o->rho != s->set.rho || o->sigma != s->set.sigma || o->scaling != s->set.scaling;
    if (o->check_termination != s->set.check_termination || o->max_iter != s->set.max_iter || refactor)
        s->retile_at = 0;                   // the learnt re-tiling point is a multiple of the old check interval
    s->set = *o;
    if (refactor) s->is_setup = false;    // like osqp_update_rho: the cached factorisation is stale
    return 0;
}

// validate per-stage state boxes [(N+1)][2][nx] and upload them (clipped to +-OSQP_INFTY) into *dst
static int upload_stage_boxes(mpcb_solver* s, const double* xbox_host, void** dst, int* has_inf, rt_stream st) {
    const size_t n = (size_t)(s->prob.horizon + 1) * 2 * s->prob.nx;
    for (int k = 0; k <= s->prob.horizon; ++k)
        for (int i = 0; i < s->prob.nx; ++i)
            if (xbox_host[(k * 2) * s->prob.nx + i] > xbox_host[(k * 2 + 1) * s->prob.nx + i])
                return fail(MPCB_E_ARG, "lower bound must be lower than or equal to upper bound");
    *has_inf = 0;
    for (size_t i = 0; i < n; ++i) if (is_big(xbox_host[i])) *has_inf = 1;
    if (!*dst) if (int r = rt_malloc(dst, n * s->esz)) return r;
    void* tmp = std::malloc(n * s->esz);
    if (!tmp) return fail(MPCB_E_ALLOC, "out of host memory");
    for (size_t i = 0; i < n; ++i) {
        if (s->esz == 4) ((float*)tmp)[i] = (float)clip_infty(xbox_host[i]);
        else ((double*)tmp)[i] = clip_infty(xbox_host[i]);
    }
    int r = rt_h2d(*dst, tmp, n * s->esz, st);
    if (!r) r = rt_sync(st);
    std::free(tmp);
    return r;
}

int mpcb_set_stage_bounds(mpcb_solver* s, const double* xbox_host) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!xbox_host) {
        rt_sync(0);
        rt_free(s->xbox); s->xbox = nullptr; s->inf_box = 0;
    } else {
        int has_inf = 0;
        if (int r = upload_stage_boxes(s, xbox_host, &s->xbox, &has_inf, 0)) return r;
        s->inf_box = has_inf;
    }
    s->inf_bounds = s->inf_prob | s->inf_box;
    s->is_setup = false;      // problem data of the next prob.setup(); after a setup use mpcb_update_bounds
    return 0;
}

size_t mpcb_workspace_bytes(const mpcb_solver* s) { return s ? s->ws_bytes : 0; }
int mpcb_num_variables(const mpcb_solver* s) { return s ? s->nvar : 0; }
int mpcb_num_constraints(const mpcb_solver* s) { return s ? s->ncon : 0; }

int mpcb_setup(mpcb_solver* s, int batch, size_t ld, const void* Ad, const void* Bd, const void* gd,
               const void* x_init, const void* Xr, void* stream) {
    if (!s || !Ad || !Bd || !x_init || !Xr) return fail(MPCB_E_ARG, "null argument");
    if (batch <= 0 || batch > s->cap) return fail(MPCB_E_ARG, "batch must be in [1, capacity]");
    if (ld != s->ld) return fail(MPCB_E_ARG, "ld must equal the solver's leading dimension (capacity rounded up to 32)");
    s->batch = batch; s->Ad = Ad; s->Bd = Bd; s->gd = gd; s->x_init = x_init; s->Xr = Xr;
    rt_stream st = (rt_stream)stream;
    int rc = dispatch(s, [&](auto* tp, auto* lp) {
        typedef typename std::remove_pointer<decltype(tp)>::type T;
        typedef typename std::remove_pointer<decltype(lp)>::type L;
        KParams<T> p = make_params<T>(s);
        if (int r = rt_memset(s->status, 0, s->ld * sizeof(int), st)) return r;
        if (int r = launch_qp<ScaleOp, T, L>(p, st)) return r;
        return launch_qp<FactorOp, T, L>(p, st);
    });
    if (rc) return rc;
    s->is_setup = true;
    s->cold_pending = true;      // osqp_setup leaves x = z = y = 0: nothing of an earlier problem may warm-start this one
    return 0;
}

int mpcb_update(mpcb_solver* s, const void* x_init, const void* Xr) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!s->is_setup) return fail(MPCB_E_STATE, "update before setup");
    if (x_init) s->x_init = x_init;
    if (Xr) s->Xr = Xr;
    return 0;
}

// osqp_update_bounds (osqp.c) + update_rho_vec (auxil.c) for one QP: did a bound row change its type between the old and
// the new bounds?  The rows' rho is a function of the scaled bounds (row_rho), evaluated on the fly by every kernel, so
// the cached factor must be rebuilt exactly when OSQP rebuilds its KKT matrix.
template <typename T, typename L>
MPCB_HD bool bounds_change_row_types(const KParams<T>& po, const KParams<T>& pn, int b) {
    constexpr int NX = L::NX, NU = L::NU;
    Ws<T, L> ws(pn, b);
    const T rho = clamp_rho(pn.rho), rho_eq = (T)kRhoEqOverRhoIneq * rho;
    bool changed = false;
    for (int k = 0; k <= pn.N; ++k) {
        const T* R = ws.R(k);
        T lo0[NX], hi0[NX], lo1[NX], hi1[NX];
        stage_box<T, L>(po, k, lo0, hi0);
        stage_box<T, L>(pn, k, lo1, hi1);
        for (int j = 0; j < NX; ++j) {
            const T E = MPCB_AT(R, L::R_E + L::OBX + j);
            changed |= row_rho(E * lo0[j], E * hi0[j], rho, rho_eq) != row_rho(E * lo1[j], E * hi1[j], rho, rho_eq);
        }
        if (k < pn.N)
            for (int j = 0; j < NU; ++j) {
                const T E = MPCB_AT(R, L::R_E + L::OBU + j);
                changed |= row_rho(E * po.umin[j], E * po.umax[j], rho, rho_eq) !=
                           row_rho(E * pn.umin[j], E * pn.umax[j], rho, rho_eq);
            }
    }
    return changed;
}
template <typename T, typename L>
struct RefactorFn {
    KParams<T> po, pn;
    MPCB_HD void operator()(int b) const {
        if (bounds_change_row_types<T, L>(po, pn, b)) factor_one<T, L>(pn, b);
    }
};

int mpcb_update_bounds(mpcb_solver* s, const double* xmin, const double* xmax, const double* umin, const double* umax,
                       const double* xbox_host, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    rt_stream st = (rt_stream)stream;
    mpcb_problem np = s->prob;
    for (int i = 0; i < np.nx; ++i) { if (xmin) np.xmin[i] = xmin[i]; if (xmax) np.xmax[i] = xmax[i]; }
    for (int i = 0; i < np.nu; ++i) { if (umin) np.umin[i] = umin[i]; if (umax) np.umax[i] = umax[i]; }
    for (int i = 0; i < np.nx; ++i)
        if (np.xmin[i] > np.xmax[i]) return fail(MPCB_E_ARG, "lower bound must be lower than or equal to upper bound");
    for (int i = 0; i < np.nu; ++i)
        if (np.umin[i] > np.umax[i]) return fail(MPCB_E_ARG, "lower bound must be lower than or equal to upper bound");
    if (!s->is_setup) {           // before prob.setup(): plain problem data
        if (xbox_host) if (int r = mpcb_set_stage_bounds(s, xbox_host)) return r;
        s->prob = np;
        s->inf_prob = problem_has_inf_bounds(np);
        s->inf_bounds = s->inf_prob | s->inf_box;
        return 0;
    }
    void* new_box = s->xbox;
    int new_inf_box = s->inf_box;
    if (xbox_host) {
        if (int r = upload_stage_boxes(s, xbox_host, &s->xbox_alt, &new_inf_box, st)) return r;
        new_box = s->xbox_alt;
    }
    int rc = dispatch(s, [&](auto* tp, auto* lp) {
        typedef typename std::remove_pointer<decltype(tp)>::type T;
        typedef typename std::remove_pointer<decltype(lp)>::type L;
        RefactorFn<T, L> fn;
        fn.po = make_params<T>(s);
        const mpcb_problem keep = s->prob;
        void* keep_box = s->xbox;
        s->prob = np; s->xbox = new_box;
        fn.pn = make_params<T>(s);
        s->prob = keep; s->xbox = keep_box;
        // either set of bounds may hold an infinity: both evaluations take the general row_rho
        return launch_1d(fn.pn.B, st, fn);
    });
    if (rc) return rc;
    s->prob = np;
    if (xbox_host) { s->xbox_alt = s->xbox; s->xbox = new_box; s->inf_box = new_inf_box; }
    s->inf_prob = problem_has_inf_bounds(np);
    s->inf_bounds = s->inf_prob | s->inf_box;
    return 0;
}

// The ADMM launch: warp-per-tile with TMA-staged stage records when two record buffers per warp fit in
// shared memory (every shape of the reference does), else one lane per QP straight from global memory.
// MPCB_NO_TMA=1 forces the latter (used to cross-check the two kernels in tests).
template <typename T, typename L>
static int launch_admm(const KParams<T>& p, mpcb_solver* s, rt_stream st) {
#ifndef MPCB_EMU
    const bool no_tma = g_opt_tma.load() == 0;
    const int max_smem = s->dev_max_smem, sms = s->dev_sms;      // of the solver's own device (queried at mpcb_create)
    const size_t per_warp = 2 * (size_t)L::REC * TILE * sizeof(T) + 16;      // two record buffers + two mbarriers
    int warps = (int)(((size_t)max_smem - 128 - 16) / per_warp);
    if (warps > 8) warps = 8;
    if (!no_tma && warps >= 2) {
        const int warps_max = warps;
        RT_CHECK(cudaFuncSetAttribute(admm_tma_kernel<T, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)max_smem - 128)));
        const int ntiles = (p.B + TILE - 1) / TILE;
        // few tiles (small batches, the straggler launch after a re-tiling): spread them over the SMs with fewer
        // warps per CTA instead of packing them on a handful of SMs — such launches are latency-bound
        int wpc = (ntiles + sms - 1) / sms;
        if (wpc > warps) wpc = warps;
        warps = wpc < 1 ? 1 : wpc;
        int grid = (ntiles + warps - 1) / warps;
        if (grid > sms) grid = sms;              // persistent CTAs, one per SM; work items are handed out dynamically
        // behind the buffers: a slice per warp for the per-QP reference (used when it does not cost a warp of shared
        // memory at full occupancy; the slot is laid out either way) and 16 bytes for the CTA's round counter
        KParams<T> pk = p;
        const size_t xr_bytes = (size_t)L::NX * TILE * sizeof(T);
        pk.xr_smem = (!p.xr_tv && (size_t)warps_max * (per_warp + xr_bytes) + 16 + 128 <= (size_t)max_smem) ? 1 : 0;
        const size_t smem = (size_t)warps * per_warp + (pk.xr_smem ? (size_t)warps * xr_bytes : 0) + 16;
        if (int r = rt_memset(s->tile_counter, 0, sizeof(int), st)) return r;
        if (int r = rt_memset(s->tile_prog, 0, (size_t)ntiles * sizeof(int), st)) return r;
        admm_tma_kernel<T, L><<<grid, warps * 32, smem, st>>>(pk, s->tile_counter);
        ++g_launches;
        return rt_launch_check("admm_tma");
    }
#endif
    (void)s;
    return launch_qp<AdmmOp, T, L>(p, st);
}

#ifdef MPCB_EMU
constexpr int WIDE_G_HOST = 0;       // tests/emu has no warp shuffles: the wide kernel does not exist there
#else
constexpr int WIDE_G_HOST = WIDE_G;
#endif
// Steady-state iterations it0+1 .. it_stop of a small, re-tiled set with 8 lanes per QP (admm_wide.cuh).  Returns 1 when
// the shape / problem flavour is not covered (the caller then lets admm_tma_kernel run those iterations), -1 on error.
template <typename T, typename L>
static int launch_wide(const KParams<T>& p, rt_stream st) {
#ifndef MPCB_EMU
    if constexpr (L::NW <= WIDE_G) {
        // worth it while the set is small: ~2400 QPs fill the GPU at 53 us per iteration (45 QP-iterations/us beyond that);
        // the main kernel needs 104 us per iteration up to ~28000 QPs — the two cross near 4700 QPs
        // (per-stage models are not staged by the main kernel — its model loads are exposed latency — so there the wide
        // kernel wins up to much larger sets)
        if (g_opt_wide.load() == 0 || p.it0 < 1 || p.B > (p.tv ? 16384 : 4608)) return 1;
        const int threads = 128, per_cta = threads / WIDE_G;
        if (p.tv) admm_wide_kernel<T, L, true><<<(p.B + per_cta - 1) / per_cta, threads, 0, st>>>(p);
        else admm_wide_kernel<T, L, false><<<(p.B + per_cta - 1) / per_cta, threads, 0, st>>>(p);
        ++g_launches;
        return rt_launch_check("admm_wide") ? -1 : 0;
    }
#endif
    (void)p; (void)st;
    return 1;
}

// copy the workspace columns of the surviving QPs from the home workspace into dense tiles of the scratch one
// (records and headers; the duals y are not needed: unsolved rows are in p-form)
template <typename T>
static int retile_impl(mpcb_solver* s, int n, const int* list, rt_stream st) {
    const size_t S1 = (size_t)(s->prob.horizon + 1), REC = (size_t)s->REC, HDR = (size_t)s->HDR;
    const T* rec = (const T*)s->rec; const T* hdr = (const T*)s->hdr;
    T* rec2 = (T*)s->rec2; T* hdr2 = (T*)s->hdr2;
    const int per = (int)(S1 * REC + HDR);
    // thread = (element, destination slot) with the slot fastest: destination writes are coalesced
    return launch_1d(n * per, st, MPCB_LAMBDA(int idx) {
        const int e = idx / n, j = idx - e * n;
        const int b = list[j];
        const size_t ts = (size_t)(b >> 5), ls = (size_t)(b & 31), td = (size_t)(j >> 5), ldn = (size_t)(j & 31);
        if ((size_t)e < S1 * REC) rec2[(td * S1 * REC + e) * TILE + ldn] = rec[(ts * S1 * REC + e) * TILE + ls];
        else { const size_t h = (size_t)e - S1 * REC; hdr2[(td * HDR + h) * TILE + ldn] = hdr[(ts * HDR + h) * TILE + ls]; }
    });
}
// bring x, z, y of the re-tiled QPs back to their home columns
template <typename T>
static int untile_impl(mpcb_solver* s, int n, const int* list, rt_stream st) {
    const size_t S1 = (size_t)(s->prob.horizon + 1), REC = (size_t)s->REC, HDR = (size_t)s->HDR, CS = (size_t)s->CS,
                 VS = (size_t)s->VS;
    const size_t R_X = VS + CS + (size_t)s->LT, nxp = VS + CS, NX = (size_t)s->prob.nx;
    T* rec = (T*)s->rec; T* hdr = (T*)s->hdr; T* yr = (T*)s->yrows;
    const T* rec2 = (const T*)s->rec2; const T* hdr2 = (const T*)s->hdr2; const T* yr2 = (const T*)s->yrows2;
    const int per = (int)(S1 * (nxp + CS) + 2 * NX);
    return launch_1d(n * per, st, MPCB_LAMBDA(int idx) {
        const int e = idx / n, j = idx - e * n;
        const int b = list[j];
        const size_t td = (size_t)(b >> 5), ldn = (size_t)(b & 31), ts = (size_t)(j >> 5), ls = (size_t)(j & 31);
        size_t r = (size_t)e;
        if (r < S1 * nxp) {                      // x and p (= z) blocks of every record
            const size_t k = r / nxp, o = R_X + (r - k * nxp);
            rec[((td * S1 + k) * REC + o) * TILE + ldn] = rec2[((ts * S1 + k) * REC + o) * TILE + ls];
            return;
        }
        r -= S1 * nxp;
        if (r < S1 * CS) { yr[(td * S1 * CS + r) * TILE + ldn] = yr2[(ts * S1 * CS + r) * TILE + ls]; return; }
        r -= S1 * CS;                             // header: p (= z) and y of the dyn_0 rows
        hdr[(td * HDR + NX + r) * TILE + ldn] = hdr2[(ts * HDR + NX + r) * TILE + ls];
    });
}

// The ADMM loop on the host side.  The device runs it in chunks of `check_termination` iterations (a chunk ends right
// after a termination test; unsolved rows stay in p-form, so chunking does not change a single bit).  After a chunk
// the number of unsolved QPs is read back; once at most half of the current set is left they are RE-TILED — their
// workspace columns are copied into dense tiles of a scratch workspace — so that warps stop streaming the records of
// 32 QPs for the sake of one straggler.  At the end the re-tiled QPs are copied back to their home columns.
// MPCB_NO_RETILE=1 (or "wide" off for a small batch) runs the whole loop as a single asynchronous launch; every other
// schedule reads the number of unsolved QPs back after each tested launch (a stream synchronisation).
static int run_admm(mpcb_solver* s, int max_iter, int check_every, int warm, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!s->is_setup) return fail(MPCB_E_STATE, "solve before setup (or settings changed since setup)");
    rt_stream st = (rt_stream)stream;
    const bool no_retile = g_opt_retile.load() == 0;
    const int retile_min = g_opt_retile_min.load();
    const int B = s->batch;
    if (s->cold_pending) { warm = 0; s->cold_pending = false; }      // first launch after prob.setup(): x = z = y = 0
    // time-varying sets up to 16384 QPs run their steady-state iterations with 8 lanes per QP (see launch_wide) — when
    // that kernel covers the shape (nx + nu <= 8); otherwise they are chunked and re-tiled like everything else
    const bool tv_wide = s->prob.time_varying && B <= 16384 && g_opt_wide.load() != 0 &&
                         s->prob.nx + s->prob.nu <= WIDE_G_HOST;
    const bool chunked = !no_retile && check_every > 0 && check_every < max_iter && B >= retile_min && !tv_wide;
    int* status = s->status;
    if (int r = launch_1d(B, st, MPCB_LAMBDA(int b) { status[b] = status[b] == -7 ? -7 : (int)kUnsolved; })) return r;
    if (int r = rt_memset(s->n_surv, 0, sizeof(int), st)) return r;
    return dispatch(s, [&](auto* tp, auto* lp) {
        typedef typename std::remove_pointer<decltype(tp)>::type T;
        typedef typename std::remove_pointer<decltype(lp)>::type L;
        KParams<T> p = make_params<T>(s);
        p.max_iter = max_iter; p.check_every = check_every; p.warm = warm;
        p.chunk_len = check_every;
        // Small batches (and time-varying sets, see launch_wide) are latency-bound from the first iteration on — a few warps
        // of the main kernel, each walking 42 dependent stage sweeps per iteration: when the 8-lanes-per-QP kernel covers
        // the shape it runs every iteration between termination tests (all_wide); iteration 1 (rows enter as explicit
        // (z, y)) and the tested iterations go through the main kernel.
        const bool all_wide = !chunked && !no_retile && check_every > 1 && check_every < max_iter && L::NW <= WIDE_G_HOST &&
                              g_opt_wide.load() != 0;
        if (!chunked && !all_wide) {
            p.it0 = 0; p.it_stop = max_iter; p.list_survivors = 0;
            return launch_admm<T, L>(p, s, st);
        }
        // Large batches: phase 1 on the home workspace up to the iteration count at which the previous solve of this
        // solver re-tiled (one launch; unknown on the first solve: explore check by check).  Either way the unsolved
        // count is read after every tested launch; once at most half of the set is left it is re-tiled into dense
        // tiles of the scratch workspace, where the stragglers finish (8 lanes per QP while the set is small enough).
        int it0 = 0, n_cur = B, which = 0;
        bool in_scratch = false;
        const int* scratch_map = nullptr;
        const bool trace = std::getenv("MPCB_TRACE") != nullptr;
        while (it0 < max_iter) {
            int stop = it0 + check_every;
            if (in_scratch) stop = max_iter;
            else if (it0 == 0 && s->retile_at > 0 && !all_wide) stop = s->retile_at;
            p.B = n_cur; p.survivors = s->surv[which]; p.qp_map = in_scratch ? scratch_map : nullptr;
            if (in_scratch || all_wide) {
                // The iterations before the next termination test run with 8 lanes per QP (the last of them also saves
                // the old state), the tested one in the main kernel.
                int next_test = (it0 / check_every + 1) * check_every;
                if (next_test > max_iter) next_test = max_iter;
                if (all_wide) stop = next_test;
                if (it0 == 0 && next_test > 1) {          // iteration 1 alone (only reached with all_wide)
                    p.it0 = 0; p.it_stop = 1; p.list_survivors = 0;
                    if (int r = launch_admm<T, L>(p, s, st)) return r;
                    it0 = 1;
                }
                if (next_test - 1 > it0) {
                    p.it0 = it0; p.it_stop = next_test - 1;
                    const int rw = launch_wide<T, L>(p, st);
                    if (rw < 0) return (int)MPCB_E_CUDA;
                    if (trace) std::fprintf(stderr, "[mpcb] wide %d..%d n=%d -> %d\n", p.it0 + 1, p.it_stop, n_cur, rw);
                    if (rw == 0) { it0 = next_test - 1; stop = next_test; }
                }
            }
            p.it0 = it0; p.it_stop = stop < max_iter ? stop : max_iter; p.list_survivors = 1;
            if (int r = launch_admm<T, L>(p, s, st)) return r;
            if (trace) std::fprintf(stderr, "[mpcb] admm %d..%d n=%d scratch=%d\n", p.it0 + 1, p.it_stop, n_cur, (int)in_scratch);
            it0 = p.it_stop;
            int n_unc = 0;
            if (int r = rt_d2h(&n_unc, s->n_surv, sizeof(int), st)) return r;
            if (int r = rt_sync(st)) return r;
            if (int r = rt_memset(s->n_surv, 0, sizeof(int), st)) return r;
            if (n_unc == 0 || it0 >= max_iter) break;
            // (the 8-lanes-per-QP kernel is latency-bound up to ~2400 QPs: compacting a smaller set buys nothing)
            if (!in_scratch && 2 * n_unc <= n_cur && (!all_wide || n_cur > 2400)) {
                if (!all_wide) s->retile_at = it0;
                // re-tile: survivors (listed by QP index = home slot) -> dense tiles of the scratch workspace
                const size_t ld2 = ((size_t)n_unc + 31) / 32 * 32, S1 = (size_t)(s->prob.horizon + 1), e = s->esz;
                if (ld2 > s->ld2) {
                    rt_free(s->rec2); rt_free(s->hdr2); rt_free(s->yrows2);
                    s->rec2 = s->hdr2 = s->yrows2 = nullptr; s->ld2 = 0;
                    const size_t want = ((size_t)s->ld / 2 + 31) / 32 * 32 > ld2 ? ((size_t)s->ld / 2 + 31) / 32 * 32 : ld2;
                    if (int r = rt_malloc(&s->rec2, S1 * s->REC * want * e)) return r;
                    if (int r = rt_malloc(&s->hdr2, (size_t)s->HDR * want * e)) return r;
                    if (int r = rt_malloc(&s->yrows2, S1 * s->CS * want * e)) return r;
                    s->ld2 = want;
                    s->ws_bytes += (S1 * s->REC + s->HDR + S1 * s->CS) * want * e;
                }
                if (int r = retile_impl<T>(s, n_unc, s->surv[which], st)) return r;
                scratch_map = s->surv[which];
                which ^= 1;                       // the next launch lists its survivors in the other buffer
                in_scratch = true;
                n_cur = n_unc;
                p.rec = (T*)s->rec2; p.hdr = (T*)s->hdr2; p.yrows = (T*)s->yrows2;
            } else if (!in_scratch && !all_wide && it0 == s->retile_at) {
                s->retile_at = 0;                 // the learnt point no longer fits this workload: explore again next time
            }
        }
        if (in_scratch)
            if (int r = untile_impl<T>(s, n_cur, scratch_map, st)) return r;
        return 0;
    });
}

int mpcb_solve(mpcb_solver* s, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    return run_admm(s, s->set.max_iter, s->set.check_termination, s->set.warm_start, stream);
}

int mpcb_iterate(mpcb_solver* s, int iters, void* stream) {
    if (iters <= 0) return fail(MPCB_E_ARG, "iters must be positive");
    return run_admm(s, iters, 0, 1, stream);
}

int mpcb_cold_start(mpcb_solver* s, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!s->is_setup) return fail(MPCB_E_STATE, "cold_start before setup");
    rt_stream st = (rt_stream)stream;
    return dispatch(s, [&](auto* tp, auto* lp) {
        typedef typename std::remove_pointer<decltype(tp)>::type T;
        typedef typename std::remove_pointer<decltype(lp)>::type L;
        KParams<T> p = make_params<T>(s);
        s->cold_pending = false;
        return launch_qp<ColdOp, T, L>(p, st);
    });
}

// ---- gather: scaled iterates of the tiled workspace -> unscaled batch-major outputs in reference order
template <typename T>
static int gather_impl(mpcb_solver* s, void* x_out, void* y_out, void* u_out, rt_stream st) {
    const int N = s->prob.horizon, nx = s->prob.nx, nu = s->prob.nu, ns = s->prob.slack ? nx : 0;
    const int VS = s->VS, CS = s->CS, REC = s->REC, nvar = s->nvar, ncon = s->ncon, B = s->batch;
    const int R_D = 0, R_E = VS, R_X = VS + CS + s->LT, H_E0 = 0, H_Y0 = 2 * nx, H_C = 3 * nx, HDR = s->HDR;
    const size_t S1 = (size_t)(N + 1);
    const T* rec = (const T*)s->rec; const T* hdr = (const T*)s->hdr; const T* yr = (const T*)s->yrows;
    // a QP that ended with an infeasibility certificate has no solution: NaN like OSQP's results (osqp.c: store_solution)
    const int* stt = s->status;
    const T nanv = (T)NAN;
#define MPCB_NO_SOLUTION(b) (stt[b] == kPrimalInfeasible || stt[b] == kDualInfeasible || \
                             stt[b] == kPrimalInfeasibleInaccurate || stt[b] == kDualInfeasibleInaccurate)
    if (x_out || u_out) {
        T* xo = (T*)x_out; T* uo = (T*)u_out;
        const int per = (N + 1) * VS;
        // one thread per (QP, internal element); consecutive threads walk the elements of one QP so
        // the batch-major writes are contiguous
        int rc = launch_1d(B * per, st, MPCB_LAMBDA(int idx) {
            const int b = idx / per, e = idx - b * per;
            const int k = e / VS, o = e - k * VS;
            const T* R = rec + (((size_t)(b >> 5) * S1 + k) * REC) * TILE + (b & 31);
            const T v = MPCB_NO_SOLUTION(b) ? nanv : R[(size_t)(R_D + o) * TILE] * R[(size_t)(R_X + o) * TILE];
            if (o < nx) { if (xo) xo[(size_t)b * nvar + k * nx + o] = v; }
            else if (o < nx + ns) { if (xo) xo[(size_t)b * nvar + (N + 1) * nx + N * nu + k * nx + (o - nx)] = v; }
            else if (k < N) {
                const int j = o - nx - ns;
                if (xo) xo[(size_t)b * nvar + (N + 1) * nx + k * nu + j] = v;
                if (uo) uo[(size_t)b * N * nu + k * nu + j] = v;
            }
        });
        if (rc) return rc;
    }
    if (y_out) {
        T* yo = (T*)y_out;
        const int per = nx + (N + 1) * CS;
        int rc = launch_1d(B * per, st, MPCB_LAMBDA(int idx) {
            const int b = idx / per, e = idx - b * per;
            const size_t tile = (size_t)(b >> 5), lane = (size_t)(b & 31);
            const T* H = hdr + tile * HDR * TILE + lane;
            const T cinv = MPCB_NO_SOLUTION(b) ? nanv : (T)1 / H[(size_t)H_C * TILE];
            if (e < nx) {                                  // rows dyn_0 live in the header
                yo[(size_t)b * ncon + e] = H[(size_t)(H_E0 + e) * TILE] * H[(size_t)(H_Y0 + e) * TILE] * cinv;
                return;
            }
            const int k = (e - nx) / CS, o = (e - nx) - k * CS;
            const T v = rec[((tile * S1 + k) * REC + R_E + o) * TILE + lane] * yr[((tile * S1 + k) * CS + o) * TILE + lane] * cinv;
            if (o < nx) { if (k < N) yo[(size_t)b * ncon + (k + 1) * nx + o] = v; }            // dyn_{k+1}
            else if (o < 2 * nx) yo[(size_t)b * ncon + (N + 1) * nx + k * nx + (o - nx)] = v;     // bx_k
            else if (k < N) yo[(size_t)b * ncon + 2 * (N + 1) * nx + k * nu + (o - 2 * nx)] = v; // bu_k
        });
        if (rc) return rc;
    }
#undef MPCB_NO_SOLUTION
    return 0;
}

int mpcb_get_solution(mpcb_solver* s, void* x_out, void* y_out, void* u_out, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!s->is_setup) return fail(MPCB_E_STATE, "get_solution before setup");
    rt_stream st = (rt_stream)stream;
    if (s->cold_pending) if (int r = mpcb_cold_start(s, stream)) return r;      // set up, not solved yet: the zeros of osqp_setup
    return s->esz == 4 ? gather_impl<float>(s, x_out, y_out, u_out, st) : gather_impl<double>(s, x_out, y_out, u_out, st);
}

int mpcb_get_info(mpcb_solver* s, int* iter_out, int* status_out, void* pri_out, void* dua_out, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    rt_stream st = (rt_stream)stream;
    const int B = s->batch;
    const int* it = s->iter; const int* stt = s->status;
    if (iter_out || status_out) {
        int rc = launch_1d(B, st, MPCB_LAMBDA(int b) {
            if (iter_out) iter_out[b] = it[b];
            if (status_out) status_out[b] = stt[b];
        });
        if (rc) return rc;
    }
    if (pri_out || dua_out) {
        const size_t e = s->esz;
        const char* pr = (const char*)s->pri; const char* du = (const char*)s->dua;
        char* po = (char*)pri_out; char* dq = (char*)dua_out;
        int rc = launch_1d(B, st, MPCB_LAMBDA(int b) {
            for (size_t i = 0; i < e; ++i) {
                if (po) po[b * e + i] = pr[b * e + i];
                if (dq) dq[b * e + i] = du[b * e + i];
            }
        });
        if (rc) return rc;
    }
    return 0;
}

// ---- layout helpers ---------------------------------------------------------------------------
template <typename T>
static int to_em(int B, int elems, size_t ld, const T* src, T* dst, rt_stream st) {
    // thread per (element, QP) with the QP index fastest: coalesced element-major writes
    return launch_1d(B * elems, st, MPCB_LAMBDA(int idx) {
        const int e = idx / B, b = idx - e * B;
        dst[(size_t)e * ld + b] = src[(size_t)b * elems + e];
    });
}
template <typename T>
static int to_bm(int B, int elems, size_t ld, const T* src, T* dst, rt_stream st) {
    return launch_1d(B * elems, st, MPCB_LAMBDA(int idx) {
        const int b = idx / elems, e = idx - b * elems;
        dst[(size_t)b * elems + e] = src[(size_t)e * ld + b];
    });
}

int mpcb_to_element_major(int dtype, int batch, int elems, size_t ld, const void* src, void* dst, void* stream) {
    if (!src || !dst || batch <= 0 || elems <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad layout arguments");
    rt_stream st = (rt_stream)stream;
    return dtype == MPCB_F32 ? to_em<float>(batch, elems, ld, (const float*)src, (float*)dst, st)
                             : to_em<double>(batch, elems, ld, (const double*)src, (double*)dst, st);
}
int mpcb_to_batch_major(int dtype, int batch, int elems, size_t ld, const void* src, void* dst, void* stream) {
    if (!src || !dst || batch <= 0 || elems <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad layout arguments");
    rt_stream st = (rt_stream)stream;
    return dtype == MPCB_F32 ? to_bm<float>(batch, elems, ld, (const float*)src, (float*)dst, st)
                             : to_bm<double>(batch, elems, ld, (const double*)src, (double*)dst, st);
}

// ---- model kernels ----------------------------------------------------------------------------
int mpcb_lateral_discretize(int dtype, int batch, size_t ld, const void* speed, const double* q, void* Ad, void* Bd,
                            void* stream) {
    if (!speed || !q || !Ad || !Bd || batch <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        LateralParams<float> lp{(float)q[0], (float)q[1], (float)q[2], (float)q[3], (float)q[4], (float)q[5], (float)q[6]};
        const float* sp = (const float*)speed; float* A = (float*)Ad; float* Bm = (float*)Bd;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { lateral_one<float>(lp, sp, A, Bm, ld, b); });
    }
    LateralParams<double> lp{q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
    const double* sp = (const double*)speed; double* A = (double*)Ad; double* Bm = (double*)Bd;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { lateral_one<double>(lp, sp, A, Bm, ld, b); });
}

int mpcb_dynamics_linearize(int dtype, int batch, size_t ld, const void* x, const void* u, const double* q, void* Ad,
                            void* Bd, void* gd, void* stream) {
    if (!x || !u || !q || !Ad || !Bd || !gd || batch <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        DynParams<float> dp{(float)q[0], (float)q[1], (float)q[2], (float)q[3], (float)q[4], (float)q[5], (float)q[6], (float)q[7]};
        const float* xx = (const float*)x; const float* uu = (const float*)u;
        float* A = (float*)Ad; float* Bm = (float*)Bd; float* g = (float*)gd;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { dynamics_one<float>(dp, xx, uu, A, Bm, g, ld, b); });
    }
    DynParams<double> dp{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]};
    const double* xx = (const double*)x; const double* uu = (const double*)u;
    double* A = (double*)Ad; double* Bm = (double*)Bd; double* g = (double*)gd;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { dynamics_one<double>(dp, xx, uu, A, Bm, g, ld, b); });
}

int mpcb_kinematics_linearize(int dtype, int batch, size_t ld, const void* x, const void* u, const double* q, void* Ad,
                              void* Bd, void* gd, void* stream) {
    if (!x || !u || !q || !Ad || !Bd || !gd || batch <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        const float wb = (float)q[0], dt = (float)q[1];
        const float* xx = (const float*)x; const float* uu = (const float*)u;
        float* A = (float*)Ad; float* Bm = (float*)Bd; float* g = (float*)gd;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { kinematics_one<float>(wb, dt, xx, uu, A, Bm, g, ld, b); });
    }
    const double wb = q[0], dt = q[1];
    const double* xx = (const double*)x; const double* uu = (const double*)u;
    double* A = (double*)Ad; double* Bm = (double*)Bd; double* g = (double*)gd;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { kinematics_one<double>(wb, dt, xx, uu, A, Bm, g, ld, b); });
}

int mpcb_dynamics_step(int dtype, int batch, size_t ld, const void* x, const void* u, const double* q, void* x_next,
                       void* alpha, void* stream) {
    if (!x || !u || !q || !x_next || batch <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        DynParams<float> dp{(float)q[0], (float)q[1], (float)q[2], (float)q[3], (float)q[4], (float)q[5], (float)q[6], (float)q[7]};
        const float* xx = (const float*)x; const float* uu = (const float*)u;
        float* xn = (float*)x_next; float* al = (float*)alpha;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { dynamics_step_one<float>(dp, xx, uu, xn, al, ld, b); });
    }
    DynParams<double> dp{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]};
    const double* xx = (const double*)x; const double* uu = (const double*)u;
    double* xn = (double*)x_next; double* al = (double*)alpha;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { dynamics_step_one<double>(dp, xx, uu, xn, al, ld, b); });
}

int mpcb_kinematics_step(int dtype, int batch, size_t ld, const void* x, const void* u, const double* q, void* x_next,
                         void* stream) {
    if (!x || !u || !q || !x_next || batch <= 0 || ld < (size_t)batch) return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        const float wb = (float)q[0], dt = (float)q[1];
        const float* xx = (const float*)x; const float* uu = (const float*)u; float* xn = (float*)x_next;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { kinematics_step_one<float>(wb, dt, xx, uu, xn, ld, b); });
    }
    const double wb = q[0], dt = q[1];
    const double* xx = (const double*)x; const double* uu = (const double*)u; double* xn = (double*)x_next;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { kinematics_step_one<double>(wb, dt, xx, uu, xn, ld, b); });
}

int mpcb_augment_increment(int dtype, int batch, size_t ld, int nx, int nu, int stages, const void* Ad, const void* Bd,
                           const void* gd, void* At, void* Bt, void* gt, void* stream) {
    if (!Ad || !Bd || !At || !Bt || batch <= 0 || nx <= 0 || nu <= 0 || stages <= 0 || ld < (size_t)batch)
        return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32) {
        const float* A = (const float*)Ad; const float* Bm = (const float*)Bd; const float* g = (const float*)gd;
        float* A2 = (float*)At; float* B2 = (float*)Bt; float* g2 = (float*)gt;
        return launch_1d(batch, st, MPCB_LAMBDA(int b) { augment_one<float>(nx, nu, stages, A, Bm, g, A2, B2, g2, ld, b); });
    }
    const double* A = (const double*)Ad; const double* Bm = (const double*)Bd; const double* g = (const double*)gd;
    double* A2 = (double*)At; double* B2 = (double*)Bt; double* g2 = (double*)gt;
    return launch_1d(batch, st, MPCB_LAMBDA(int b) { augment_one<double>(nx, nu, stages, A, Bm, g, A2, B2, g2, ld, b); });
}

template <typename T>
static int plant_impl(int B, size_t ld, int nx, int nu, int shared, const T* A, const T* Bm, const T* g, const T* x,
                      const T* u, int us, T* xn, rt_stream st) {
    return launch_1d(B, st, MPCB_LAMBDA(int b) {
        const size_t mld = shared ? 1 : ld, bo = shared ? 0 : (size_t)b;
        T xo[MAXNX], acc[MAXNX];
        for (int j = 0; j < nx; ++j) xo[j] = x[(size_t)j * ld + b];
        for (int i = 0; i < nx; ++i) {
            T v = g ? g[(size_t)i * mld + bo] : (T)0;
            for (int j = 0; j < nx; ++j) v += A[(size_t)(i * nx + j) * mld + bo] * xo[j];
            for (int j = 0; j < nu; ++j) v += Bm[(size_t)(i * nu + j) * mld + bo] * u[(size_t)b * us + j];
            acc[i] = v;
        }
        for (int i = 0; i < nx; ++i) xn[(size_t)i * ld + b] = acc[i];
    });
}

int mpcb_plant_step(int dtype, int batch, size_t ld, int nx, int nu, int shared_model, const void* A, const void* Bm,
                    const void* g, const void* x, const void* u, int u_stride, void* x_next, void* stream) {
    if (!A || !Bm || !x || !u || !x_next || batch <= 0 || nx <= 0 || nx > MAXNX || nu <= 0 || nu > MAXNU ||
        ld < (size_t)batch || u_stride < nu)
        return fail(MPCB_E_ARG, "bad arguments");
    rt_stream st = (rt_stream)stream;
    if (dtype == MPCB_F32)
        return plant_impl<float>(batch, ld, nx, nu, shared_model, (const float*)A, (const float*)Bm, (const float*)g,
                                 (const float*)x, (const float*)u, u_stride, (float*)x_next, st);
    return plant_impl<double>(batch, ld, nx, nu, shared_model, (const double*)A, (const double*)Bm, (const double*)g,
                              (const double*)x, (const double*)u, u_stride, (double*)x_next, st);
}

// ---- explicit QP --------------------------------------------------------------------------------
int mpcb_qp_pattern(const mpcb_solver* s, int* Ap, int* Ai) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    const int N = s->prob.horizon, nx = s->prob.nx, nu = s->prob.nu, ns = s->prob.slack ? nx : 0;
    const int bx0 = (N + 1) * nx, bu0 = 2 * (N + 1) * nx;
    int nnz = 0, col = 0;
    for (int k = 0; k <= N; ++k)
        for (int j = 0; j < nx; ++j) {
            if (Ap) Ap[col] = nnz;
            if (Ai) Ai[nnz] = k * nx + j;
            ++nnz;
            if (k < N)
                for (int i = 0; i < nx; ++i) { if (Ai) Ai[nnz] = (k + 1) * nx + i; ++nnz; }
            if (Ai) Ai[nnz] = bx0 + k * nx + j;
            ++nnz; ++col;
        }
    for (int k = 0; k < N; ++k)
        for (int j = 0; j < nu; ++j) {
            if (Ap) Ap[col] = nnz;
            for (int i = 0; i < nx; ++i) { if (Ai) Ai[nnz] = (k + 1) * nx + i; ++nnz; }
            if (Ai) Ai[nnz] = bu0 + k * nu + j;
            ++nnz; ++col;
        }
    for (int k = 0; k <= N && ns; ++k)
        for (int j = 0; j < nx; ++j) {
            if (Ap) Ap[col] = nnz;
            if (Ai) Ai[nnz] = bx0 + k * nx + j;
            ++nnz; ++col;
        }
    if (Ap) Ap[col] = nnz;
    return nnz;
}

int mpcb_build_qp(mpcb_solver* s, void* Pdiag, void* q, void* Avals, void* l, void* u, void* stream) {
    if (!s) return fail(MPCB_E_ARG, "null solver");
    if (!s->Ad || !s->x_init) return fail(MPCB_E_STATE, "build_qp before setup");
    rt_stream st = (rt_stream)stream;
    BuildOut o{Pdiag, q, Avals, l, u};
    return dispatch(s, [&](auto* tp, auto* lp) {
        typedef typename std::remove_pointer<decltype(tp)>::type T;
        typedef typename std::remove_pointer<decltype(lp)>::type L;
        KParams<T> p = make_params<T>(s);
        return launch_1d(p.B, st, BuildFn<T, L>{p, o});
    });
}

// ---- host front door ----------------------------------------------------------------------------
static int ensure(void** p, size_t* have, size_t need) {
    if (*have >= need) return 0;
    rt_free(*p); *p = nullptr; *have = 0;
    if (rt_malloc(p, need)) return MPCB_E_ALLOC;
    *have = need;
    return 0;
}

int mpcb_solve_host(mpcb_solver* s, int batch, const void* Ad, const void* Bd, const void* gd, const void* x_init,
                    const void* Xr, void* x_out, void* u_out, int* iter_out, int* status_out) {
    if (!s || !Ad || !Bd || !x_init || !Xr) return fail(MPCB_E_ARG, "null argument");
    if (batch <= 0 || batch > s->cap) return fail(MPCB_E_ARG, "batch must be in [1, capacity]");
    const mpcb_problem& q = s->prob;
    const int N = q.horizon, nx = q.nx, nu = q.nu;
    const size_t e = s->esz, ld = s->ld;
    const int stages = q.time_varying ? N : 1;
    const int nA = stages * nx * nx, nB = stages * nx * nu, nG = gd ? stages * nx : 0, nX0 = nx,
              nXr = (q.stage_reference ? N + 1 : 1) * nx;
    const int mb = q.shared_model ? 1 : batch;                // model rows on the host side
    const size_t model_elems = (size_t)nA + nB + nG;
    const size_t in_bytes = (model_elems * mb + (size_t)(nX0 + nXr) * batch) * e;
    const size_t mld = q.shared_model ? 1 : ld;
    const size_t soa_bytes = (model_elems * mld + (size_t)(nX0 + nXr) * ld) * e;
    const size_t out_elems = (size_t)(x_out ? s->nvar : 0) + (u_out ? N * nu : 0);
    const size_t out_bytes = out_elems * batch * e + 2 * (size_t)batch * sizeof(int);
    if (ensure(&s->stage_in, &s->stage_in_bytes, in_bytes) || ensure(&s->soa_in, &s->soa_in_bytes, soa_bytes) ||
        ensure(&s->stage_out, &s->stage_out_bytes, out_bytes))
        return MPCB_E_ALLOC;
    rt_stream st = 0;
    char* din = (char*)s->stage_in;
    char* soa = (char*)s->soa_in;
    struct Part { const void* h; int elems; int rows; size_t pld; const void** dev; };
    const void *dA = nullptr, *dB = nullptr, *dG = nullptr, *dX0 = nullptr, *dXr = nullptr;
    Part parts[] = {{Ad, nA, mb, mld, &dA}, {Bd, nB, mb, mld, &dB}, {gd, nG, mb, mld, &dG},
                    {x_init, nX0, batch, ld, &dX0}, {Xr, nXr, batch, ld, &dXr}};
    for (auto& pt : parts) {
        if (!pt.h || pt.elems == 0) continue;
        const size_t nbytes = (size_t)pt.elems * pt.rows * e;
        if (int r = rt_h2d(din, pt.h, nbytes, st)) return r;
        if (pt.rows == 1 && pt.pld == 1) {
            *pt.dev = din;                                  // a single row is already element-major
        } else {
            if (int r = mpcb_to_element_major(q.dtype, pt.rows, pt.elems, pt.pld, din, soa, st)) return r;
            *pt.dev = soa;
            soa += (size_t)pt.elems * pt.pld * e;
        }
        din += nbytes;
    }
    if (int r = mpcb_setup(s, batch, ld, dA, dB, dG, dX0, dXr, st)) return r;
    if (int r = run_admm(s, s->set.max_iter, s->set.check_termination, 0, st)) return r;
    char* dout = (char*)s->stage_out;
    void* dx = nullptr; void* du = nullptr;
    if (x_out) { dx = dout; dout += (size_t)s->nvar * batch * e; }
    if (u_out) { du = dout; dout += (size_t)N * nu * batch * e; }
    int* dit = (int*)dout; int* dst = dit + batch;
    if (int r = mpcb_get_solution(s, dx, nullptr, du, st)) return r;
    if (int r = mpcb_get_info(s, dit, dst, nullptr, nullptr, st)) return r;
    if (x_out) if (int r = rt_d2h(x_out, dx, (size_t)s->nvar * batch * e, st)) return r;
    if (u_out) if (int r = rt_d2h(u_out, du, (size_t)N * nu * batch * e, st)) return r;
    if (iter_out) if (int r = rt_d2h(iter_out, dit, (size_t)batch * sizeof(int), st)) return r;
    if (status_out) if (int r = rt_d2h(status_out, dst, (size_t)batch * sizeof(int), st)) return r;
    return rt_sync(st);
}

