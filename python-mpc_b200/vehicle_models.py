"""Host-side mirror of /root/reference/Vehicle_Dynamics/vehicle_models.py, batched on the GPU.

Same class names, constructor arguments and method names as the reference; every method accepts
either ONE vehicle (numpy, the reference's shapes — drop-in) or a batch (leading dimension B) and
runs the corresponding kernel of libmpc_b200.so (csrc/models.cuh).

    Vehicle_Dynamics.get_dynamics_model(x, u)      vehicle_models.py:52-340
    Vehicle_Kinematics.get_kinematics_model(x, u)  vehicle_models.py:835-863
    Vehicle_Lateral.get_lateral_model(v)           the lateral bicycle model whose discretisation the
        reference hard-codes for one speed in vehicle_lateral_mpc_slack_increment.py:32-43; here it
        is discretised per vehicle speed (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import MPCB_F32, MPCB_F64, ptr


def _dt(dtype):
    return MPCB_F32 if dtype == torch.float32 else MPCB_F64


class _ModelBase:
    def __init__(self, dtype=torch.float64, _backend=None):
        self._be = _backend
        self.dtype = dtype

    @property
    def be(self):
        if self._be is None:
            self._be = _lib.cuda_backend()
        return self._be

    def _em(self, a, elems):
        """(B, elems) batch-major -> element-major [elems, ld]; returns (tensor, B, ld)."""
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
        t = t.to(device=self.be.device, dtype=self.dtype).reshape(-1, elems).contiguous()
        B = t.shape[0]
        ld = (B + 31) // 32 * 32
        dst = torch.zeros((elems, ld), device=self.be.device, dtype=self.dtype)
        self.be.check(self.be.lib.mpcb_to_element_major(_dt(self.dtype), B, elems, ld, ptr(t), ptr(dst), self.be.stream()))
        return dst, B, ld

    def _bm(self, em, B, shape):
        elems, ld = em.shape
        dst = torch.empty((B, elems), device=self.be.device, dtype=self.dtype)
        self.be.check(self.be.lib.mpcb_to_batch_major(_dt(self.dtype), B, elems, ld, ptr(em), ptr(dst), self.be.stream()))
        return dst.reshape((B,) + shape)


class Vehicle_Lateral(_ModelBase):
    """Lateral bicycle (error) model, state [side-slip, yaw-rate, yaw-error, lateral-error], input steer.

        beta'  = -(Cf+Cr)/(m v) beta + ((Cr lr - Cf lf)/(m v^2) - 1) r + Cf/(m v) delta
        r'     = (Cr lr - Cf lf)/Iz beta - (Cf lf^2 + Cr lr^2)/(Iz v) r + Cf lf/Iz delta
        e_yaw' = r ;  e_y' = v beta + v e_yaw
    discretised exactly (zero-order hold) per speed.  The default cornering stiffnesses / inertia are
    the least-squares fit to the literals of vehicle_lateral_mpc_slack_increment.py:32-43, which this
    model reproduces to their printed precision at v = 8.31 m/s (tests/test_oracle.py::test_lateral_model_matches_reference_literals, tests/parity_cases.py::check_models_against_reference)."""

    NOMINAL_SPEED = 8.3128334

    def __init__(self, m=1300., l_f=1.25, l_r=1.40, Iz=2555.88174, Cf=11979.9261, Cr=11140.9949, dt=0.02,
                 dtype=torch.float64, _backend=None):
        super().__init__(dtype, _backend)
        self.m, self.l_f, self.l_r, self.Iz, self.Cf, self.Cr, self.dt = m, l_f, l_r, Iz, Cf, Cr, dt
        self.nx, self.nu = 4, 1

    def _params(self):
        return (C.c_double * 7)(self.m, self.l_f, self.l_r, self.Iz, self.Cf, self.Cr, self.dt)

    def lateral_model_em(self, speed_em, B, ld, out=None):
        """speed [ld] device tensor -> element-major Ad [16, ld], Bd [4, ld] (the hot-path form)."""
        if out is not None:
            Ad, Bd = out
        else:
            Ad = torch.empty((16, ld), device=self.be.device, dtype=self.dtype)
            Bd = torch.empty((4, ld), device=self.be.device, dtype=self.dtype)
        self.be.check(self.be.lib.mpcb_lateral_discretize(_dt(self.dtype), B, ld, ptr(speed_em), self._params(), ptr(Ad),
                                                          ptr(Bd), self.be.stream()))
        return Ad, Bd

    def get_lateral_model(self, v):
        """v scalar -> (Ad (4,4), Bd (4,1)) numpy;  v (B,) -> torch tensors (B,4,4), (B,4,1)."""
        single = np.ndim(v) == 0
        sp, B, ld = self._em(np.atleast_1d(np.asarray(v, dtype=np.float64)) if not isinstance(v, torch.Tensor) else v, 1)
        Ad, Bd = self.lateral_model_em(sp, B, ld)
        A, Bm = self._bm(Ad, B, (4, 4)), self._bm(Bd, B, (4, 1))
        if single:
            return A[0].cpu().numpy(), Bm[0].cpu().numpy()
        return A, Bm


class Vehicle_Dynamics(_ModelBase):
    """Same constructor as the reference (vehicle_models.py:27-50)."""

    def __init__(self, m=1300, l_f=1.25, l_r=1.40, width=1.78, length=4.25, turning_circle=10.4, C_d=0.34, A_f=2.0,
                 C_roll=0.015, dt=0.02, dtype=torch.float64, _backend=None):
        super().__init__(dtype, _backend)
        self.m, self.l_f, self.l_r = m, l_f, l_r
        self.wheelbase = l_f + l_r
        self.width, self.length, self.turning_circle = width, length, turning_circle
        self.max_steer = math.atan(self.wheelbase / turning_circle)
        self.Iz = 1 / 12 * m * (width ** 2 + length ** 2)
        self.C_d, self.A_f, self.C_roll, self.roh, self.dt = C_d, A_f, C_roll, 1.23, dt
        self.nx, self.nu = 6, 2

    def _params(self):
        return (C.c_double * 8)(self.m, self.l_f, self.l_r, self.Iz, self.C_d, self.A_f, self.C_roll, self.dt)

    def dynamics_model_em(self, x_em, u_em, B, ld):
        mk = lambda n: torch.empty((n, ld), device=self.be.device, dtype=self.dtype)
        Ad, Bd, gd = mk(36), mk(12), mk(6)
        self.be.check(self.be.lib.mpcb_dynamics_linearize(_dt(self.dtype), B, ld, ptr(x_em), ptr(u_em), self._params(),
                                                          ptr(Ad), ptr(Bd), ptr(gd), self.be.stream()))
        return Ad, Bd, gd

    def get_dynamics_model(self, x, u):
        """Single vehicle: x (6,)|(6,1), u (2,)|(2,1) -> numpy Ad (6,6), Bd (6,2), gd (6,1) like the reference.
        Batch: x (B,6), u (B,2) -> torch (B,6,6), (B,6,2), (B,6)."""
        xa = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
        single = (xa.ndim == 1) or (xa.ndim == 2 and xa.shape[1] == 1 and xa.shape[0] == 6)
        if single:
            xa = np.asarray(xa, dtype=np.float64).reshape(1, 6)
            u = np.asarray(u, dtype=np.float64).reshape(1, 2)
        x_em, B, ld = self._em(xa, 6)
        u_em, _, _ = self._em(u, 2)
        Ad, Bd, gd = self.dynamics_model_em(x_em, u_em, B, ld)
        A, Bm, g = self._bm(Ad, B, (6, 6)), self._bm(Bd, B, (6, 2)), self._bm(gd, B, (6,))
        if single:
            return A[0].cpu().numpy(), Bm[0].cpu().numpy(), g[0].cpu().numpy().reshape(6, 1)
        return A, Bm, g


    def update_dynamics_model(self, x, u):
        """The nonlinear plant step (vehicle_models.py:343-482).  Single vehicle: returns (x_next (6,1), alpha_f, alpha_r)
        like the reference.  Batch: x (B,6), u (B,2) -> torch x_next (B,6), alpha (B,2)."""
        xa = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
        single = (xa.ndim == 1) or (xa.ndim == 2 and xa.shape[1] == 1 and xa.shape[0] == 6)
        if single:
            xa = np.asarray(xa, dtype=np.float64).reshape(1, 6)
            u = np.asarray(u, dtype=np.float64).reshape(1, 2)
        x_em, B, ld = self._em(xa, 6)
        u_em, _, _ = self._em(u, 2)
        xn = torch.empty((6, ld), device=self.be.device, dtype=self.dtype)
        al = torch.empty((2, ld), device=self.be.device, dtype=self.dtype)
        self.be.check(self.be.lib.mpcb_dynamics_step(_dt(self.dtype), B, ld, ptr(x_em), ptr(u_em), self._params(), ptr(xn),
                                                     ptr(al), self.be.stream()))
        xn, al = self._bm(xn, B, (6,)), self._bm(al, B, (2,))
        if single:
            a = al[0].cpu().numpy()
            return xn[0].cpu().numpy().reshape(6, 1), float(a[0]), float(a[1])
        return xn, al


class Vehicle_Kinematics(_ModelBase):
    """Same constructor as the reference (vehicle_models.py:829-833)."""

    def __init__(self, l_f=1.25, l_r=1.40, dt=0.02, dtype=torch.float64, _backend=None):
        super().__init__(dtype, _backend)
        self.l_f, self.l_r, self.wheelbase, self.dt = l_f, l_r, l_f + l_r, dt
        self.nx, self.nu = 4, 2

    def kinematics_model_em(self, x_em, u_em, B, ld):
        mk = lambda n: torch.empty((n, ld), device=self.be.device, dtype=self.dtype)
        A, Bm, Cv = mk(16), mk(8), mk(4)
        par = (C.c_double * 2)(self.wheelbase, self.dt)
        self.be.check(self.be.lib.mpcb_kinematics_linearize(_dt(self.dtype), B, ld, ptr(x_em), ptr(u_em), par, ptr(A),
                                                            ptr(Bm), ptr(Cv), self.be.stream()))
        return A, Bm, Cv

    def get_kinematics_model(self, x, u):
        xa = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
        single = (xa.ndim == 1) or (xa.ndim == 2 and xa.shape[1] == 1 and xa.shape[0] == 4)
        if single:
            xa = np.asarray(xa, dtype=np.float64).reshape(1, 4)
            u = np.asarray(u, dtype=np.float64).reshape(1, 2)
        x_em, B, ld = self._em(xa, 4)
        u_em, _, _ = self._em(u, 2)
        A, Bm, Cv = self.kinematics_model_em(x_em, u_em, B, ld)
        A, Bm, Cv = self._bm(A, B, (4, 4)), self._bm(Bm, B, (4, 2)), self._bm(Cv, B, (4,))
        if single:
            return A[0].cpu().numpy(), Bm[0].cpu().numpy(), Cv[0].cpu().numpy().reshape(4, 1)
        return A, Bm, Cv


    def update_kinematics_model(self, x, u):
        """vehicle_models.py:866-882.  Single vehicle: x (4,) -> x_next (4,) (the reference mutates and returns x; here a
        new array).  Batch: x (B,4), u (B,2) -> torch (B,4)."""
        xa = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
        single = (xa.ndim == 1) or (xa.ndim == 2 and xa.shape[1] == 1 and xa.shape[0] == 4)
        shape = None if not single else np.asarray(xa).shape
        if single:
            xa = np.asarray(xa, dtype=np.float64).reshape(1, 4)
            u = np.asarray(u, dtype=np.float64).reshape(1, 2)
        x_em, B, ld = self._em(xa, 4)
        u_em, _, _ = self._em(u, 2)
        xn = torch.empty((4, ld), device=self.be.device, dtype=self.dtype)
        par = (C.c_double * 2)(self.wheelbase, self.dt)
        self.be.check(self.be.lib.mpcb_kinematics_step(_dt(self.dtype), B, ld, ptr(x_em), ptr(u_em), par, ptr(xn),
                                                       self.be.stream()))
        xn = self._bm(xn, B, (4,))
        return xn[0].cpu().numpy().reshape(shape) if single else xn


def augment_increment_em(be, dtype, Ad, Bd, gd, B, ld, nx, nu, stages=1, out=None):
    """delta-u augmentation on element-major model arrays (mpc_dynamics.py:337-341)."""
    na = nx + nu
    mk = lambda n: torch.empty((n, ld), device=be.device, dtype=dtype)
    if out is not None:
        At, Bt = out
    else:
        At, Bt = mk(stages * na * na), mk(stages * na * nu)
    gt = mk(stages * na) if gd is not None else None
    be.check(be.lib.mpcb_augment_increment(_dt(dtype), B, ld, nx, nu, stages, ptr(Ad), ptr(Bd), ptr(gd), ptr(At), ptr(Bt),
                                           ptr(gt), be.stream()))
    return At, Bt, gt
