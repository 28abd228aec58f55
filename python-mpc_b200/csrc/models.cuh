// Batched vehicle-model kernels of the QP build: one thread per QP, element-major I/O.
//
//   lateral_one     lateral bicycle model, ZOH-discretised per vehicle speed.  The reference
//                   hard-codes this model's (Ad_sys, Bd_sys) for one speed in
//                   vehicle_lateral_mpc_slack_increment.py:32-43; with the default parameters
//                   (python-mpc_b200/vehicle_models.py: Vehicle_Lateral) and v = 8.31 m/s this
//                   function reproduces those literals to their printed precision.
//   dynamics_one    Vehicle_Dynamics.get_dynamics_model  (Vehicle_Dynamics/vehicle_models.py:52-340):
//                   Pacejka lateral tyres, analytic Jacobians, forward-Euler discretisation.
//   kinematics_one  Vehicle_Kinematics.get_kinematics_model (vehicle_models.py:835-863).
//   augment_one     delta-u augmentation (mpc_dynamics.py:337-341).
#pragma once
#include "mpc_common.h"

namespace mpcb {

template <typename T>
struct LateralParams { T m, lf, lr, Iz, Cf, Cr, dt; };

// exp(M) for a small dense K x K matrix: scaling and squaring around a degree-12 Taylor polynomial
template <typename T, int K>
MPCB_HD void expm_small(T (&M)[K][K], T (&Eo)[K][K]) {
    T nrm = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        T r = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) r += (M[i][j] < 0 ? -M[i][j] : M[i][j]);
        nrm = r > nrm ? r : nrm;
    }
    int s = 0;
    T scale = 1;
    while (nrm * scale > (T)0.25 && s < 40) { scale *= (T)0.5; ++s; }
    T X[K][K], term[K][K], tmp[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            X[i][j] = M[i][j] * scale;
            term[i][j] = (i == j) ? (T)1 : (T)0;
            Eo[i][j] = term[i][j];
        }
    for (int q = 1; q <= 12; ++q) {
        const T inv = (T)1 / (T)q;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                T acc = 0;
#pragma unroll
                for (int l = 0; l < K; ++l) acc += term[i][l] * X[l][j];
                tmp[i][j] = acc * inv;
            }
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) { term[i][j] = tmp[i][j]; Eo[i][j] += tmp[i][j]; }
    }
    for (int q = 0; q < s; ++q) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                T acc = 0;
#pragma unroll
                for (int l = 0; l < K; ++l) acc += Eo[i][l] * Eo[l][j];
                tmp[i][j] = acc;
            }
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) Eo[i][j] = tmp[i][j];
    }
}

// state [beta, yaw_rate, e_yaw, e_y], input steer
template <typename T>
MPCB_HD void lateral_one(const LateralParams<T>& q, const T* speed, T* Ad, T* Bd, size_t ld, int b) {
    T v = speed[b];
    const T vmin = (T)0.1;
    if (v >= 0 && v < vmin) v = vmin;
    if (v < 0 && v > -vmin) v = -vmin;
    T M[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) M[i][j] = 0;
    const T dt = q.dt;
    M[0][0] = -(q.Cf + q.Cr) / (q.m * v) * dt;
    M[0][1] = ((q.Cr * q.lr - q.Cf * q.lf) / (q.m * v * v) - (T)1) * dt;
    M[1][0] = (q.Cr * q.lr - q.Cf * q.lf) / q.Iz * dt;
    M[1][1] = -(q.Cf * q.lf * q.lf + q.Cr * q.lr * q.lr) / (q.Iz * v) * dt;
    M[2][1] = dt;
    M[3][0] = v * dt;
    M[3][2] = v * dt;
    M[0][4] = q.Cf / (q.m * v) * dt;
    M[1][4] = q.Cf * q.lf / q.Iz * dt;
    T Eo[5][5];
    expm_small<T, 5>(M, Eo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Ad[(size_t)(i * 4 + j) * ld + b] = Eo[i][j];
        Bd[(size_t)i * ld + b] = Eo[i][4];
    }
}

template <typename T>
struct DynParams { T m, lf, lr, Iz, Cd, Af, Croll, dt; };

MPCB_HD float msin(float v) { return sinf(v); }
MPCB_HD double msin(double v) { return sin(v); }
MPCB_HD float mcos(float v) { return cosf(v); }
MPCB_HD double mcos(double v) { return cos(v); }
MPCB_HD float matan(float v) { return atanf(v); }
MPCB_HD double matan(double v) { return atan(v); }
MPCB_HD float matan2(float a, float c) { return atan2f(a, c); }
MPCB_HD double matan2(double a, double c) { return atan2(a, c); }
MPCB_HD float mtan(float v) { return tanf(v); }
MPCB_HD double mtan(double v) { return tan(v); }

template <typename T>
MPCB_HD void dynamics_one(const DynParams<T>& q, const T* xg, const T* ug, T* Ad, T* Bd, T* gd, size_t ld, int b) {
    const T m = q.m, lf = q.lf, lr = q.lr, Iz = q.Iz, dt = q.dt, roh = (T)1.23;
    const T wb = lf + lr;
    // Pacejka lateral tyre constants (vehicle_models.py:114-132)
    const T a0 = (T)-22.1, a1 = (T)1011, a2 = (T)1078, a3 = (T)1.82, a4 = (T)0.208;
    const T Clat = (T)1.30, r2d = (T)(180.0 / 3.14159265358979323846);
    const T Fzf = (T)9.81 * (m * lr / wb) * (T)0.001, Fzr = (T)9.81 * (m * lf / wb) * (T)0.001;
    const T Df = a0 * Fzf * Fzf + a1 * Fzf, Dr = a0 * Fzr * Fzr + a1 * Fzr;
    const T Bf = a2 * msin(a3 * matan(a4 * Fzf)) / (Clat * Df) * r2d;
    const T Br = a2 * msin(a3 * matan(a4 * Fzr)) / (Clat * Dr) * r2d;
    T x[6], u[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xg[(size_t)i * ld + b];
    u[0] = ug[b]; u[1] = ug[ld + b];
    // low-speed guards (vehicle_models.py:143-159)
    if (x[3] >= 0 && x[3] < (T)0.5) { x[4] = 0; x[5] = 0; u[0] = 0; if (x[3] < (T)0.3) x[3] = (T)0.3; }
    if (x[3] > (T)-0.5 && x[3] < 0) { x[4] = 0; x[5] = 0; u[0] = 0; if (x[3] > (T)-0.3) x[3] = (T)-0.3; }
    const T yaw = x[2], vx = x[3], vy = x[4], wz = x[5], st = u[0], acc = u[1];
    const T af = -matan2(lf * wz + vy, vx) + st, ar = -matan2(-lr * wz + vy, vx);
    const T Fyf = Df * msin(Clat * matan(Bf * af)), Fyr = Dr * msin(Clat * matan(Br * ar));
    const T sg = vx > 0 ? (T)1 : (vx < 0 ? (T)-1 : (T)0);
    const T Rroll = q.Croll * m * (T)9.81 * sg, Faero = (T)0.5 * roh * q.Cd * q.Af * vx * vx * sg;
    const T Fxf = m * acc - Faero - Rroll;
    const T sy = msin(yaw), cy = mcos(yaw), ss = msin(st), cs = mcos(st);
    T f[6];
    f[0] = vx * cy - vy * sy;
    f[1] = vy * cy + vx * sy;
    f[2] = wz;
    f[3] = (T)1 / m * (Fxf * cs - Fyf * ss + m * vy * wz);
    f[4] = (T)1 / m * (Fxf * ss + Fyr + Fyf * cs - m * vx * wz);
    f[5] = (T)1 / Iz * (Fxf * lf * ss + Fyf * lf * cs - Fyr * lr);
    const T dFx_dvx = -roh * q.Cd * q.Af * vx, dFx_da = m;
    const T kf = (Bf * Clat * Df * mcos(Clat * matan(Bf * af))) / ((T)1 + Bf * Bf * af * af);
    const T kr = (Br * Clat * Dr * mcos(Clat * matan(Br * ar))) / ((T)1 + Br * Br * ar * ar);
    const T nf = lf * wz + vy, nr = -lr * wz + vy;
    const T df_den = nf * nf + vx * vx, dr_den = nr * nr + vx * vx;
    const T dFyf_dvx = kf * nf / df_den, dFyf_dvy = kf * (-vx / df_den), dFyf_dw = kf * (-lf * vx) / df_den, dFyf_ds = kf;
    const T dFyr_dvx = kr * nr / dr_den, dFyr_dvy = kr * (-vx) / dr_den, dFyr_dw = kr * (lr * vx) / dr_den;
    T Ac[6][6], Bc[6][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) Ac[i][j] = 0;
        Bc[i][0] = 0; Bc[i][1] = 0;
    }
    Ac[0][2] = -vx * sy - vy * cy; Ac[0][3] = cy; Ac[0][4] = -sy;
    Ac[1][2] = -vy * sy + vx * cy; Ac[1][3] = sy; Ac[1][4] = cy;
    Ac[2][5] = 1;
    Ac[3][3] = (T)1 / m * (dFx_dvx * cs - dFyf_dvx * ss);
    Ac[3][4] = (T)1 / m * (-dFyf_dvy * ss + m * wz);
    Ac[3][5] = (T)1 / m * (-dFyf_dw * ss + m * vy);
    Ac[4][3] = (T)1 / m * (dFx_dvx * ss + dFyr_dvx + dFyf_dvx * cs - m * wz);
    Ac[4][4] = (T)1 / m * (dFyr_dvy + dFyf_dvy * cs);
    Ac[4][5] = (T)1 / m * (dFyr_dw + dFyf_dw * cs - m * vx);
    Ac[5][3] = (T)1 / Iz * (dFx_dvx * lf * ss + dFyf_dvx * lf * cs - dFyr_dvx * lr);
    Ac[5][4] = (T)1 / Iz * (dFyf_dvy * lf * cs - dFyr_dvy * lr);
    Ac[5][5] = (T)1 / Iz * (dFyf_dw * lf * cs - dFyr_dw * lr);
    Bc[3][0] = (T)1 / m * (-Fxf * ss - dFyf_ds * ss - Fyf * cs);
    Bc[3][1] = (T)1 / m * (dFx_da * cs);
    Bc[4][0] = (T)1 / m * (Fxf * cs + dFyf_ds * cs - Fyf * ss);
    Bc[4][1] = (T)1 / m * (dFx_da * ss);
    Bc[5][0] = (T)1 / Iz * (Fxf * lf * cs + dFyf_ds * lf * cs - Fyf * lf * ss);
    Bc[5][1] = (T)1 / Iz * (dFx_da * lf * ss);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        T gc = f[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) gc -= Ac[i][j] * x[j];
        gc -= Bc[i][0] * u[0] + Bc[i][1] * u[1];
#pragma unroll
        for (int j = 0; j < 6; ++j) Ad[(size_t)(i * 6 + j) * ld + b] = (i == j ? (T)1 : (T)0) + Ac[i][j] * dt;
        Bd[(size_t)(i * 2 + 0) * ld + b] = Bc[i][0] * dt;
        Bd[(size_t)(i * 2 + 1) * ld + b] = Bc[i][1] * dt;
        gd[(size_t)i * ld + b] = gc * dt;
    }
}

// The nonlinear plant step of the same model (Vehicle_Dynamics.update_dynamics_model, vehicle_models.py:343-482): one
// explicit Euler step with the Pacejka lateral tyre forces; also returns the axle side-slip angles.  Used by the
// reference's scripts for the simulated vehicle and for extending the prediction by one stage (mpc_dynamics.py:607).
template <typename T>
MPCB_HD void dynamics_step_one(const DynParams<T>& q, const T* xg, const T* ug, T* xn, T* alpha, size_t ld, int b) {
    const T m = q.m, lf = q.lf, lr = q.lr, Iz = q.Iz, dt = q.dt, roh = (T)1.23;
    const T wb = lf + lr;
    const T a0 = (T)-22.1, a1 = (T)1011, a2 = (T)1078, a3 = (T)1.82, a4 = (T)0.208;
    const T Clat = (T)1.30, r2d = (T)(180.0 / 3.14159265358979323846);
    const T Fzf = (T)9.81 * (m * lr / wb) * (T)0.001, Fzr = (T)9.81 * (m * lf / wb) * (T)0.001;
    const T Df = a0 * Fzf * Fzf + a1 * Fzf, Dr = a0 * Fzr * Fzr + a1 * Fzr;
    const T Bf = a2 * msin(a3 * matan(a4 * Fzf)) / (Clat * Df) * r2d;
    const T Br = a2 * msin(a3 * matan(a4 * Fzr)) / (Clat * Dr) * r2d;
    T x[6], u[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xg[(size_t)i * ld + b];
    u[0] = ug[b]; u[1] = ug[ld + b];
    // low-speed guards (vehicle_models.py:416-433); like the reference they modify the state the step starts from
    if (x[3] >= 0 && x[3] < (T)0.5) { x[4] = 0; x[5] = 0; u[0] = 0; if (x[3] < (T)0.3) x[3] = (T)0.3; }
    if (x[3] > (T)-0.5 && x[3] < 0) { x[4] = 0; x[5] = 0; u[0] = 0; if (x[3] > (T)-0.3) x[3] = (T)-0.3; }
    const T yaw = x[2], vx = x[3], vy = x[4], wz = x[5], st = u[0], acc = u[1];
    const T af = -matan2(lf * wz + vy, vx) + st, ar = -matan2(-lr * wz + vy, vx);
    const T Fyf = Df * msin(Clat * matan(Bf * af)), Fyr = Dr * msin(Clat * matan(Br * ar));
    const T sg = vx > 0 ? (T)1 : (vx < 0 ? (T)-1 : (T)0);
    const T Rroll = q.Croll * m * (T)9.81 * sg, Faero = (T)0.5 * roh * q.Cd * q.Af * vx * vx * sg;
    const T Fxf = m * acc - Faero - Rroll;
    const T sy = msin(yaw), cy = mcos(yaw), ss = msin(st), cs = mcos(st);
    T f[6];
    f[0] = vx * cy - vy * sy;
    f[1] = vy * cy + vx * sy;
    f[2] = wz;
    f[3] = (T)1 / m * (Fxf * cs - Fyf * ss + m * vy * wz);
    f[4] = (T)1 / m * (Fxf * ss + Fyr + Fyf * cs - m * vx * wz);
    f[5] = (T)1 / Iz * (Fxf * lf * ss + Fyf * lf * cs - Fyr * lr);
#pragma unroll
    for (int i = 0; i < 6; ++i) xn[(size_t)i * ld + b] = x[i] + f[i] * dt;
    if (alpha) { alpha[b] = af; alpha[ld + b] = ar; }
}

// Vehicle_Kinematics.update_kinematics_model (vehicle_models.py:866-882).  The reference updates the state in place,
// so the yaw update already sees the NEW speed — reproduced.
template <typename T>
MPCB_HD void kinematics_step_one(T wheelbase, T dt, const T* xg, const T* ug, T* xn, size_t ld, int b) {
    const T px = xg[b], py = xg[ld + b], v = xg[2 * ld + b], yaw = xg[3 * ld + b];
    const T st = ug[b], acc = ug[ld + b];
    const T vn = v + acc * dt;
    xn[b] = px + v * mcos(yaw) * dt;
    xn[ld + b] = py + v * msin(yaw) * dt;
    xn[2 * ld + b] = vn;
    xn[3 * ld + b] = yaw + vn / wheelbase * mtan(st) * dt;
}

template <typename T>
MPCB_HD void kinematics_one(T wheelbase, T dt, const T* xg, const T* ug, T* A, T* Bm, T* C, size_t ld, int b) {
    const T v = xg[2 * ld + b], yaw = xg[3 * ld + b], st = ug[b];
    const T sy = msin(yaw), cy = mcos(yaw), ct = mcos(st);
    T Am[4][4] = {{1, 0, dt * cy, -dt * v * sy}, {0, 1, dt * sy, dt * v * cy}, {0, 0, 1, 0}, {0, 0, dt * mtan(st) / wheelbase, 1}};
    T Bv[4][2] = {{0, 0}, {0, 0}, {0, dt}, {dt * v / (wheelbase * ct * ct), 0}};
    T Cv[4] = {dt * v * sy * yaw, -dt * v * cy * yaw, 0, -dt * v * st / (wheelbase * ct * ct)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) A[(size_t)(i * 4 + j) * ld + b] = Am[i][j];
        Bm[(size_t)(i * 2) * ld + b] = Bv[i][0];
        Bm[(size_t)(i * 2 + 1) * ld + b] = Bv[i][1];
        C[(size_t)i * ld + b] = Cv[i];
    }
}

// (Ad, Bd, gd) of `stages` stages -> augmented (A~, B~, g~); runtime nx, nu
template <typename T>
MPCB_HD void augment_one(int nx, int nu, int stages, const T* Ad, const T* Bd, const T* gd, T* At, T* Bt, T* gt,
                         size_t ld, int b) {
    const int na = nx + nu;
    for (int k = 0; k < stages; ++k) {
        const size_t oa = (size_t)k * nx * nx, ob = (size_t)k * nx * nu, og = (size_t)k * nx;
        const size_t ta = (size_t)k * na * na, tb = (size_t)k * na * nu, tg = (size_t)k * na;
        for (int i = 0; i < na; ++i) {
            for (int j = 0; j < na; ++j) {
                T v;
                if (i < nx) v = j < nx ? Ad[(oa + i * nx + j) * ld + b] : Bd[(ob + i * nu + (j - nx)) * ld + b];
                else v = (j == i) ? (T)1 : (T)0;
                At[(ta + i * na + j) * ld + b] = v;
            }
            for (int j = 0; j < nu; ++j)
                Bt[(tb + i * nu + j) * ld + b] = i < nx ? Bd[(ob + i * nu + j) * ld + b] : ((i - nx) == j ? (T)1 : (T)0);
            if (gt) gt[(tg + i) * ld + b] = (i < nx && gd) ? gd[(og + i) * ld + b] : (T)0;
        }
    }
}

}  // namespace mpcb
