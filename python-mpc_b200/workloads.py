"""Synthetic workloads of the named shapes (BASELINE.json configs): seeded random vehicle speeds,
initial states and references.  Inputs only — no solver logic and no oracle here."""
from __future__ import annotations

import numpy as np
import torch

from .mpc import LateralMPC

DEG = np.pi / 180.0


class LateralWorkload:
    """Batch of lateral-MPC QPs, H = N.  Constants are those of vehicle_lateral_mpc_slack_increment.py:56-64."""

    def __init__(self, B, N, slack, increment, seed, dtype, shared_speed=None):
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
        self.x0 = x0
        self.xr = np.zeros((B, 4))
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


def lateral_vanilla_shared(B, N=20, seed=0, dtype=torch.float64, speed=8.3128334):
    """configs[1]: vanilla lateral MPC, ONE shared linearisation (one speed)."""
    return LateralWorkload(B, N, False, False, seed, dtype, shared_speed=speed)
