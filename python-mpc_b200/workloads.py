"""Synthetic workloads of the named shapes (BASELINE.json configs): seeded random vehicle speeds,
initial states and references.  Inputs only — no solver logic and no oracle here."""
from __future__ import annotations

import numpy as np
import torch

from .mpc import LateralMPC

DEG = np.pi / 180.0


class LateralWorkload:
    """Batch of lateral-MPC QPs, H = N.  Constants are those of vehicle_lateral_mpc_slack_increment.py:57-74."""

    def __init__(self, B, N, slack, increment, seed, dtype, shared_speed=None, state_scale=1.0, ref_scale=0.0):
        rng = np.random.default_rng(seed)
        self.B, self.N, self.slack, self.increment, self.dtype = B, N, slack, increment, dtype
        self.Q = np.array([5., 5., 10., 10.]); self.R = np.array([10.])
        self.W = np.array([10., 10., 10., 10., 0.]) if increment else np.array([10., 10., 10., 10.])
        self.S = np.array([1., 1., 1., 1., 0.]) if increment else np.ones(4)
        if increment:
            self.xmin = np.array([-np.pi, -0.5 * np.pi, -15 * DEG, -10., -30 * DEG]); self.xmax = -self.xmin
            self.umin = np.array([-0.5 * DEG]); self.umax = np.array([0.5 * DEG])
        else:
            self.xmin = np.array([-np.pi, -0.5 * np.pi, -15 * DEG, -10.]); self.xmax = -self.xmin
            self.umin = np.array([-30 * DEG]); self.umax = np.array([30 * DEG])
        nx = 5 if increment else 4
        self.nx = nx
        self.speed = np.full(B, shared_speed) if shared_speed is not None else rng.uniform(5.0, 30.0, B)
        x0 = np.zeros((B, nx))
        x0[:, 0] = rng.uniform(-0.05, 0.05, B)        # side slip
        x0[:, 1] = rng.uniform(-0.2, 0.2, B)          # yaw rate
        x0[:, 2] = rng.uniform(-8 * DEG, 8 * DEG, B)  # yaw error
        x0[:, 3] = rng.uniform(-4.0, 4.0, B)          # lateral error
        if increment:
            x0[:, 4] = rng.uniform(-10 * DEG, 10 * DEG, B)   # previous steer
        self.x0 = x0 * state_scale
        self.xr = np.zeros((B, 4))
        if ref_scale:
            self.xr[:, 3] = rng.uniform(-ref_scale, ref_scale, B)      # lateral-offset reference (lane position)
        self.shared_speed = shared_speed

    def make_controller(self, capacity=None, vehicle=None, _backend=None, **settings):
        kw = dict(slack=self.slack, increment=self.increment, dtype=self.dtype, capacity=capacity or self.B,
                  _backend=_backend)
        if self.slack:
            kw.update(W=self.W, S=self.S)
        if self.shared_speed is not None:
            from .vehicle_models import Vehicle_Lateral
            veh = vehicle or Vehicle_Lateral(dtype=self.dtype, _backend=_backend)
            Ad, Bd = veh.get_lateral_model(float(self.shared_speed))
            kw.update(Ad=Ad, Bd=Bd)
        return LateralMPC(self.N, self.Q, self.R, self.xmin, self.xmax, self.umin, self.umax, vehicle=vehicle, **kw,
                          **settings)


def lateral_slack_increment(B, N=20, seed=0, dtype=torch.float32):
    """configs[2]: soft-constraint + incremental lateral MPC, per-QP speed linearisation."""
    return LateralWorkload(B, N, True, True, seed, dtype)


def lateral_closed_loop_sweep(B, N=20, seed=0, dtype=torch.float64):
    """configs[4]: scenarios of a closed-loop sweep — the configs[2] controller around its reference: initial states within
    5 % of the configs[2] ranges (|e_y| <= 0.2 m, yaw error <= 0.4 deg, ...) and a random lateral-offset reference within
    +-0.1 m.  (From the full configs[2] ranges the H = 20 controller of the reference script, rate-limited to 0.5 deg per step,
    does not stabilise its own plant model: |e_y| grows past 30 m within 200 steps at most speeds and the QPs end up
    infeasible — no closed loop to sweep.)"""
    return LateralWorkload(B, N, True, True, seed, dtype, state_scale=0.05, ref_scale=0.1)


def lateral_vanilla_shared(B, N=20, seed=0, dtype=torch.float64, speed=8.3128334):
    """configs[1]: vanilla lateral MPC, ONE shared linearisation (one speed)."""
    return LateralWorkload(B, N, False, False, seed, dtype, shared_speed=speed)


class DynamicWorkload:
    """configs[3]: long-horizon MPC over the combined longitudinal-lateral dynamics model
    (Vehicle_Dynamics, nx = 6, nu = 2), linearised stage by stage along each vehicle's own predicted trajectory
    (the Ad_list / Bd_list / gd_list of Control/MPC/mpc_dynamics.py:520-527).  Weights and bounds of
    mpc_dynamics.py:445-456 (vanilla form: bounds on the inputs instead of their increments)."""

    def __init__(self, B, N=100, seed=0, dtype=torch.float64, dt=0.05):
        rng = np.random.default_rng(seed)
        self.B, self.N, self.dtype, self.dt = B, N, dtype, dt
        self.Q = np.array([100., 100., 100., 50., 50., 50.]); self.QN = 10 * self.Q; self.R = np.array([50., 50.])
        self.xmin = np.array([-np.inf, -np.inf, -2 * np.pi, -100., -30., -0.5 * np.pi]); self.xmax = -self.xmin
        self.umin = np.array([-15 * DEG, -3.]); self.umax = np.array([15 * DEG, 1.])
        x0 = np.zeros((B, 6))
        x0[:, 1] = rng.uniform(-1.0, 1.0, B)            # lateral offset from the path y = 0
        x0[:, 2] = rng.uniform(-6 * DEG, 6 * DEG, B)    # heading
        x0[:, 3] = rng.uniform(6.0, 20.0, B)            # v_x
        x0[:, 4] = rng.uniform(-0.2, 0.2, B)
        x0[:, 5] = rng.uniform(-0.1, 0.1, B)
        self.x0 = x0
        self.u0 = np.zeros((B, 2))
        self.v_ref = 10.0

    def references(self):
        """Xr (B, 6, N+1): follow the line y = 0 at v_ref, positions advancing with the initial speed."""
        B, N = self.B, self.N
        Xr = np.zeros((B, 6, N + 1))
        Xr[:, 0, :] = self.x0[:, 0:1] + self.x0[:, 3:4] * self.dt * np.arange(N + 1)[None, :]
        Xr[:, 3, :] = self.v_ref
        return Xr


def rollout_linearisation(vehicle, x0, u0, N):
    """Per-stage linearisation along the predicted trajectory x_{k+1} = A_k x_k + B_k u0 + g_k, on the device.
    Returns element-major time-varying model arrays (N*36, ld), (N*12, ld), (N*6, ld) and ld."""
    from ._lib import ptr
    from .vehicle_models import _dt
    be = vehicle.be
    x_em, B, ld = vehicle._em(x0, 6)
    u_em, _, _ = vehicle._em(u0, 2)
    u_bm = (u0 if isinstance(u0, torch.Tensor) else torch.as_tensor(np.asarray(u0))).to(be.device, vehicle.dtype).contiguous()
    mk = lambda n: torch.empty((N * n, ld), device=be.device, dtype=vehicle.dtype)
    A, Bm, g = mk(36), mk(12), mk(6)
    for k in range(N):
        Ak, Bk, gk = A[k * 36:(k + 1) * 36], Bm[k * 12:(k + 1) * 12], g[k * 6:(k + 1) * 6]
        be.check(be.lib.mpcb_dynamics_linearize(_dt(vehicle.dtype), B, ld, ptr(x_em), ptr(u_em), vehicle._params(), ptr(Ak),
                                                ptr(Bk), ptr(gk), be.stream()))
        xn = torch.empty_like(x_em)
        be.check(be.lib.mpcb_plant_step(_dt(vehicle.dtype), B, ld, 6, 2, 0, ptr(Ak), ptr(Bk), ptr(gk), ptr(x_em), ptr(u_bm), 2,
                                        ptr(xn), be.stream()))
        x_em = xn
    return A, Bm, g, B, ld
