"""Host-side mirror of the reference's MPC entry points.

    mpc(Ad_mat, Bd_mat, gd_mat, x_vec, Xr, Q, QN, R, N, xmin, xmax, umin, umax)
        Control/MPC/mpc_kinematics.py:150-213 — returns an OSQP-like result (res.x, res.info.status)
    mpc_lists(Ad_list, Bd_list, gd_list, x_vec, Xr, pred_x, pred_u, Q, QN, R, N, xmin, xmax, umin, umax)
        Control/MPC/mpc_dynamics.py:156-282 — returns (pred_x, pred_u)
    mpc_increment(Ad_list, Bd_list, gd_list, x_tilda_vec, Xr, pred_x_tilda, pred_del_u, Q, QN, R, N, ...)
        Control/MPC/mpc_dynamics.py:284-436, mpc_incre_kine_func.py:84-222 — returns (pred_x_tilda, pred_del_u)
    LateralMPC
        the controller of vehicle_lateral_mpc_slack_increment.py (vanilla / slack / delta-u / slack +
        delta-u over the lateral bicycle model) as an object: constructor(horizon, weights, bounds),
        solve(state, reference) -> input sequence, plus the batched entry point solve_batch().

Argument names, meaning, variable ordering of ``res.x`` and the error behaviour ("OSQP did not solve
the problem!") follow the reference.  All numerical work happens in libmpc_b200.so.

Deviation from the reference's defaults: the reference calls ``prob.setup(...)`` with OSQP's defaults, where
``adaptive_rho`` (and, in mpc_dynamics.py, ``polish=False``) is on; this path runs OSQP's ADMM with a FIXED rho
(north_star: adaptive_rho and polish off) and rejects ``adaptive_rho=True`` / ``polish=True``.  ``res.info.iter`` therefore
differs from a stock OSQP run, and a problem that stock OSQP solves by adapting rho may need an explicit ``rho=`` here
(the reference script's own N = 100 closed loop needs rho = 10, see oracle/make_golden.py); ``res.info.status`` and the
print / raise paths follow the status this path reaches.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .solver import BatchSolver
from .vehicle_models import Vehicle_Lateral, augment_increment_em, _dt
from ._lib import ptr


def _diag_of(M):
    if hasattr(M, "diagonal") and getattr(M, "ndim", 1) == 2:
        return np.asarray(M.diagonal(), dtype=np.float64).ravel()
    return np.asarray(M, dtype=np.float64).ravel()


class _Info:
    pass


class Result:
    """Shape of ``res`` returned by ``prob.solve()`` for one QP."""

    def __init__(self, x, y, it, status_val, pri, dua):
        self.x, self.y = x, y
        self.info = _Info()
        self.info.iter, self.info.status_val = int(it), int(status_val)
        self.info.status = _lib.STATUS_STRING.get(int(status_val), "unknown")
        self.info.pri_res, self.info.dua_res = float(pri), float(dua)


class BatchResult:
    def __init__(self, x, u, info):
        self.x, self.u, self.info = x, u, info


def _single_result(s):
    x, y, _ = s.solution(want_y=True, want_u=False)
    inf = s.info()
    return Result(x[0].cpu().numpy().astype(np.float64), y[0].cpu().numpy().astype(np.float64), inf.iter[0],
                  inf.status_val[0], inf.pri_res[0], inf.dua_res[0])


_SOLVERS = {}


def _cached_solver(key, make):
    s = _SOLVERS.get(key)
    if s is None:
        if len(_SOLVERS) > 16:
            _SOLVERS.pop(next(iter(_SOLVERS))).close()
        s = _SOLVERS[key] = make()
    return s


def _key(*parts):
    out = []
    for p in parts:
        out.append(tuple(np.asarray(p, dtype=np.float64).ravel().tolist()) if isinstance(p, (np.ndarray, list)) else p)
    return tuple(out)


def mpc(Ad_mat, Bd_mat, gd_mat, x_vec, Xr, Q, QN, R, N, xmin, xmax, umin, umax, dtype=torch.float64, _backend=None,
        **osqp_settings):
    """Vanilla MPC for one time-invariant linearisation (mpc_kinematics.py:150).  Returns ``res``."""
    Ad = np.asarray(Ad_mat, dtype=np.float64); Bd = np.asarray(Bd_mat, dtype=np.float64)
    nx, nu = Bd.shape
    settings = dict(warm_start=True); settings.update(osqp_settings)
    key = _key("mpc", N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), xmin, xmax, umin, umax, str(dtype),
               tuple(sorted(settings.items())), id(_backend))
    s = _cached_solver(key, lambda: BatchSolver(N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), xmin, xmax, umin, umax,
                                                dtype=dtype, stage_reference=True, capacity=1, _backend=_backend, **settings))
    gd = None if gd_mat is None else np.asarray(gd_mat, dtype=np.float64).reshape(1, nx)
    Xr = np.asarray(Xr, dtype=np.float64)
    if Xr.ndim == 1 or Xr.shape[1] == 1:
        Xr = np.tile(Xr.reshape(nx, 1), (1, N + 1))
    s.setup(Ad[None], Bd[None], gd, np.asarray(x_vec, dtype=np.float64).reshape(1, nx), Xr[None])
    s.cold_start()      # a fresh osqp.OSQP() object is created on every call of the reference's mpc()
    s.solve()
    return _single_result(s)


def mpc_(Ad_mat, Bd_mat, gd_mat, x_vec, Xr, Q, QN, R, N, lb_x, ub_x, lb_y, ub_y, umin, umax, dtype=torch.float64,
         _backend=None, **osqp_settings):
    """Vanilla MPC with a per-stage position corridor (mpc_kinematics.py:202, mpc_kinematics_pred_matrix.py:203): stage i
    has the state box [lb_x[i], lb_y[i], -10, -pi] .. [ub_x[i], ub_y[i], 10, pi], exactly the reference's literals.
    Returns ``res``."""
    Ad = np.asarray(Ad_mat, dtype=np.float64); Bd = np.asarray(Bd_mat, dtype=np.float64)
    nx, nu = Bd.shape
    if nx != 4 or len(lb_x) != N + 1:
        raise ValueError("mpc_ expects the kinematic model (nx = 4) and N + 1 corridor entries")
    lo = np.stack([np.array([lb_x[i], lb_y[i], -10.0, -np.pi]) for i in range(N + 1)])
    hi = np.stack([np.array([ub_x[i], ub_y[i], 10.0, np.pi]) for i in range(N + 1)])
    settings = dict(warm_start=True); settings.update(osqp_settings)
    key = _key("mpc_", N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), umin, umax, str(dtype),
               tuple(sorted(settings.items())), id(_backend))
    s = _cached_solver(key, lambda: BatchSolver(N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), lo[0], hi[0], umin, umax,
                                                dtype=dtype, stage_reference=True, capacity=1, _backend=_backend, **settings))
    s.set_stage_bounds(lo, hi)
    gd = None if gd_mat is None else np.asarray(gd_mat, dtype=np.float64).reshape(1, nx)
    s.setup(Ad[None], Bd[None], gd, np.asarray(x_vec, dtype=np.float64).reshape(1, nx), np.asarray(Xr, dtype=np.float64)[None])
    s.cold_start()
    s.solve()
    return _single_result(s)


def mpc__(Ad_list, Bd_list, gd_list, x_vec, Xr, Q, QN, R, N, xmin, xmax, umin, umax, dtype=torch.float64, _backend=None,
          **osqp_settings):
    """Vanilla MPC over the per-stage linearisations along the previous prediction ("predictive linearised matrix",
    mpc_kinematics_pred_matrix.py:268).  Same QP as mpc_lists; returns ``res`` like the reference's mpc__."""
    Ad = np.stack([np.asarray(a, dtype=np.float64) for a in Ad_list[:N]]); Bd = np.stack([np.asarray(b, dtype=np.float64) for b in Bd_list[:N]])
    gd = np.stack([np.asarray(g, dtype=np.float64).reshape(-1) for g in gd_list[:N]])
    nx, nu = Bd.shape[1:]
    settings = dict(warm_start=True); settings.update(osqp_settings)
    s = _tv_solver(N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), xmin, xmax, umin, umax, dtype, _backend, settings)
    s.setup(Ad[None], Bd[None], gd[None], np.asarray(x_vec, dtype=np.float64).reshape(1, nx), np.asarray(Xr, dtype=np.float64)[None])
    s.cold_start(); s.solve()
    return _single_result(s)


def _tv_solver(N, nx, nu, Q, QN, R, xmin, xmax, umin, umax, dtype, _backend, settings):
    key = _key("tv", N, nx, nu, Q, QN, R, xmin, xmax, umin, umax, str(dtype), tuple(sorted(settings.items())), id(_backend))
    return _cached_solver(key, lambda: BatchSolver(N, nx, nu, Q, QN, R, xmin, xmax, umin, umax, dtype=dtype,
                                                   time_varying=True, stage_reference=True, capacity=1,
                                                   _backend=_backend, **settings))


def mpc_lists(Ad_list, Bd_list, gd_list, x_vec, Xr, pred_x, pred_u, Q, QN, R, N, xmin, xmax, umin, umax,
              dtype=torch.float64, _backend=None, **osqp_settings):
    """Vanilla MPC over per-stage linearisations (mpc_dynamics.py:156).  Returns (pred_x, pred_u)."""
    # the reference reads the first N entries of the lists (its main() loops may hand over N + 1)
    Ad = np.stack([np.asarray(a, dtype=np.float64) for a in Ad_list[:N]]); Bd = np.stack([np.asarray(b, dtype=np.float64) for b in Bd_list[:N]])
    gd = np.stack([np.asarray(g, dtype=np.float64).reshape(-1) for g in gd_list[:N]])
    nx, nu = Bd.shape[1:]
    settings = dict(warm_start=True, polish=False); settings.update(osqp_settings)
    s = _tv_solver(N, nx, nu, _diag_of(Q), _diag_of(QN), _diag_of(R), xmin, xmax, umin, umax, dtype, _backend, settings)
    s.setup(Ad[None], Bd[None], gd[None], np.asarray(x_vec, dtype=np.float64).reshape(1, nx), np.asarray(Xr)[None])
    s.cold_start(); s.solve()
    res = _single_result(s)
    if res.info.status != "solved":
        print("OSQP did not solve the problem!")
    sol_state, sol_action = res.x[:-N * nu], res.x[-N * nu:]
    pred_x[:, :] = sol_state.reshape(N + 1, nx).T
    pred_u[:, :N] = sol_action.reshape(N, nu).T
    pred_u[:, -1] = pred_u[:, -2]
    return pred_x, pred_u


def mpc_increment(Ad_list, Bd_list, gd_list, x_tilda_vec, Xr, pred_x_tilda, pred_del_u, Q, QN, R, N, xmin_tilda,
                  xmax_tilda, del_umin, del_umax, dtype=torch.float64, strict=False, _backend=None, **osqp_settings):
    """Incremental (delta-u) MPC over per-stage linearisations (mpc_dynamics.py:284, mpc_incre_kine_func.py:84).
    strict=True raises like mpc_incre_kine_func.py:179-181; otherwise prints like mpc_dynamics.py:405-407."""
    # the reference reads the first N entries of the lists (its main() loops may hand over N + 1)
    Ad = np.stack([np.asarray(a, dtype=np.float64) for a in Ad_list[:N]]); Bd = np.stack([np.asarray(b, dtype=np.float64) for b in Bd_list[:N]])
    gd = np.stack([np.asarray(g, dtype=np.float64).reshape(-1) for g in gd_list[:N]])
    nx, nu = Bd.shape[1:]
    na = nx + nu
    settings = dict(warm_start=False, polish=False); settings.update(osqp_settings)
    z = np.zeros(nu)
    s = _tv_solver(N, na, nu, np.concatenate([_diag_of(Q), z]), np.concatenate([_diag_of(QN), z]), _diag_of(R), xmin_tilda,
                   xmax_tilda, del_umin, del_umax, dtype, _backend, settings)
    be = s.be
    # delta-u augmentation runs on the device (mpcb_augment_increment), per stage
    em = lambda a, e: s.to_element_major(torch.as_tensor(a).reshape(1, e), 1, e, s.ld)
    At, Bt, gt = augment_increment_em(be, dtype, em(Ad, N * nx * nx), em(Bd, N * nx * nu), em(gd, N * nx), 1, s.ld, nx, nu, N)
    Xr = np.asarray(Xr, dtype=np.float64)
    Xrt = np.vstack([Xr, np.zeros((nu, N + 1))]).T.reshape(1, -1)              # stage-major (N+1, na)
    s.batch = 1
    s.setup(At, Bt, gt, em(np.asarray(x_tilda_vec, dtype=np.float64), na), em(Xrt, (N + 1) * na), element_major=True)
    s.cold_start(); s.solve()
    res = _single_result(s)
    if res.info.status != "solved":
        if strict:
            raise ValueError("OSQP did not solve the problem!")
        print("OSQP did not solve the problem!")
    sol_state, sol_action = res.x[:-N * nu], res.x[-N * nu:]
    pred_x_tilda[:, :] = sol_state.reshape(N + 1, na).T
    pred_del_u[:, :N] = sol_action.reshape(N, nu).T
    pred_del_u[:, -1] = pred_del_u[:, -2]
    return pred_x_tilda, pred_del_u


class LateralMPC:
    """Lateral MPC controller (vehicle_lateral_mpc_slack_increment.py as an object).

    Constructor: horizon, weights, bounds.  ``solve(state, reference)`` returns the input sequence
    for one vehicle; ``solve_batch(states, references, speeds)`` solves B vehicles at once, each
    linearised at its own speed.

    formulation   state passed to solve()        bounds
    ------------  ----------------------------   -----------------------------------------------
    increment=0   x (4)                          xmin/xmax (4), umin/umax (1) on the steer
    increment=1   x~ = [x; u_prev] (5)           xmin/xmax (5) = (x_min, u_min), umin/umax on delta-u
    slack=1 adds slack variables on the state bounds with cost W and coupling S (W_tilda, weight_slack_tilda).
    """

    def __init__(self, N, Q, R, xmin, xmax, umin, umax, QN=None, slack=False, increment=False, W=None, S=None,
                 vehicle=None, Ad=None, Bd=None, dtype=torch.float64, capacity=1, _backend=None, **osqp_settings):
        self.N, self.slack, self.increment, self.dtype = int(N), bool(slack), bool(increment), dtype
        self.vehicle = vehicle if vehicle is not None else Vehicle_Lateral(dtype=dtype, _backend=_backend)
        self.nx_sys, self.nu = 4, 1
        self.nx = self.nx_sys + (self.nu if increment else 0)
        Qd = _diag_of(Q); QNd = Qd if QN is None else _diag_of(QN)
        if increment and Qd.size == self.nx_sys:      # Q_tilda = C~' Q C~ (vehicle_lateral_mpc_slack_increment.py:65-69)
            Qd = np.concatenate([Qd, np.zeros(self.nu)]); QNd = np.concatenate([QNd, np.zeros(self.nu)])
        if slack:
            W = np.ones(self.nx) if W is None else _diag_of(W)
            S = np.ones(self.nx) if S is None else _diag_of(S)
        self.shared = Ad is not None
        self.solver = BatchSolver(N, self.nx, self.nu, Qd, QNd, _diag_of(R), xmin, xmax, umin, umax, slack=slack, W=W, S=S,
                                  dtype=dtype, shared_model=self.shared, capacity=capacity, _backend=_backend,
                                  **osqp_settings)
        self.be = self.solver.be
        self._fixed = None
        if self.shared:
            self.set_model(Ad, Bd)
        self._is_setup = False

    def set_model(self, Ad, Bd):
        """One shared linearisation for every QP (e.g. the reference's hard-coded Ad_sys, Bd_sys)."""
        if not self.shared:
            raise ValueError("controller was built for per-QP speed linearisation; pass Ad/Bd to the constructor")
        dev = lambda a, e: torch.as_tensor(np.asarray(a, dtype=np.float64)).to(self.be.device, self.dtype).reshape(e, 1).contiguous()
        A, Bm = dev(Ad, self.nx_sys ** 2), dev(Bd, self.nx_sys * self.nu)
        if self.increment:
            A, Bm, _ = augment_increment_em(self.be, self.dtype, A, Bm, None, 1, 1, self.nx_sys, self.nu)
        self._fixed = (A, Bm)
        self._is_setup = False

    def _model(self, speeds, B):
        if self.shared:
            return self._fixed
        s = self.solver
        if speeds is None:
            raise ValueError("speeds are required: the controller linearises the lateral model per vehicle speed")
        sp = speeds if isinstance(speeds, torch.Tensor) else torch.as_tensor(np.asarray(speeds, dtype=np.float64))
        sp = s.to_element_major(sp.reshape(B, 1), B, 1, s.ld, out=s.buffer("speed", (1, s.ld)))
        A, Bm = self.vehicle.lateral_model_em(sp, B, s.ld, out=(s.buffer("Ad", (16, s.ld)), s.buffer("Bd", (4, s.ld))))
        if self.increment:
            na = self.nx
            A, Bm, _ = augment_increment_em(self.be, self.dtype, A, Bm, None, B, s.ld, self.nx_sys, self.nu,
                                            out=(s.buffer("At", (na * na, s.ld)), s.buffer("Bt", (na * self.nu, s.ld))))
        return A, Bm

    def _pad_ref(self, refs, B):
        r = refs if isinstance(refs, torch.Tensor) else torch.as_tensor(np.asarray(refs, dtype=np.float64))
        r = r.to(self.be.device, self.dtype).reshape(B, -1)
        if r.shape[1] == self.nx_sys and self.increment:
            r = torch.cat([r, torch.zeros((B, self.nu), device=r.device, dtype=r.dtype)], dim=1)
        return r

    def solve_batch(self, states, references, speeds=None, want_x=True, reuse=False):
        """QP build (discretise per speed, augment) + setup (scale, factor) + ADMM + gather, all on the device.
        states (B, nx), references (B, nx_sys | nx), speeds (B,).  Returns BatchResult(x (B,nvar), u (B,N,nu), info).
        reuse=True: the result tensors are views of persistent buffers (overwritten by the next call)."""
        s = self.solver
        x0 = states if isinstance(states, torch.Tensor) else torch.as_tensor(np.asarray(states, dtype=np.float64))
        x0 = x0.to(self.be.device, self.dtype).reshape(-1, self.nx)
        B = x0.shape[0]
        if B > s.capacity:
            raise ValueError("batch %d exceeds controller capacity %d" % (B, s.capacity))
        A, Bm = self._model(speeds, B)
        s.batch = B
        s.setup(A, Bm, None, s.to_element_major(x0, B, self.nx, s.ld, out=s.buffer("x0", (self.nx, s.ld))),
                s.to_element_major(self._pad_ref(references, B), B, self.nx, s.ld, out=s.buffer("xr", (self.nx, s.ld))),
                element_major=True)
        self._is_setup = True
        s.solve()
        x, _, u = s.solution(want_x=want_x, want_y=False, want_u=True, reuse=reuse)
        return BatchResult(x, u, s.info(reuse=reuse))

    def update_batch(self, states, references=None):
        """prob.update(q, l, u) + prob.solve() for a batch already set up (closed loop, warm start)."""
        s = self.solver
        if not self._is_setup:
            raise _lib.MpcError("update_batch before solve_batch")
        B = s.batch
        x0 = states if isinstance(states, torch.Tensor) else torch.as_tensor(np.asarray(states, dtype=np.float64))
        s.update(x_init=x0.to(self.be.device, self.dtype).reshape(B, self.nx),
                 Xr=None if references is None else self._pad_ref(references, B))
        s.solve()
        x, _, u = s.solution(want_x=True, want_y=False, want_u=True)
        return BatchResult(x, u, s.info())

    def update_bounds(self, xmin=None, xmax=None, umin=None, umax=None):
        """New state / input bounds for the controller already set up — the l / u part of the reference's
        prob.update(q=q_new, l=l_new, u=u_new) (vehicle_lateral_mpc_slack_increment.py:158-172, :237), e.g. the
        tightened lateral-error bound xmin_tilda[3] = 2 of steps 401..900.  Same layout as the constructor's bounds.
        Takes effect with the next solve() / update_batch(); scaling, factor (unless a row changes type) and the
        warm start are kept, like OSQP's update."""
        self.solver.update_bounds(xmin=xmin, xmax=xmax, umin=umin, umax=umax)
        return self

    def closed_loop_batch(self, states, references, speeds=None, steps=1, record=True, bounds_at=None):
        """Closed-loop sweep (BASELINE configs[4]): B scenarios, `steps` MPC steps each, everything on the device.
        Step 0 sets the QPs up (scale + factor) and solves them from a cold start; every later step is the reference's
        prob.update(l=, u=) with the new initial state followed by a warm-started prob.solve()
        (vehicle_lateral_mpc_slack_increment.py:237-257); the plant is the QP's own model, x+ = A~ x + B~ du0 (:256-257).
        bounds_at(k) -> None | dict(xmin=, xmax=, umin=, umax=): bounds that apply from step k on (the reference tightens
        xmin_tilda[3] at step 401 and relaxes it at 901, :158-172); applied through update_bounds before step k's solve.
        Returns (trajectory (steps+1, B, nx) or None, applied inputs (steps, B, nu), iterations (steps, B))."""
        s = self.solver
        if bounds_at is not None and bounds_at(0):
            self.update_bounds(**bounds_at(0))        # before setup: plain problem data
        res = self.solve_batch(states, references, speeds, want_x=False)
        B, nx, nu, N = s.batch, self.nx, self.nu, self.N
        A, Bm = s._keep["Ad"], s._keep["Bd"]
        x_em = s._keep["x_init"]
        dev = self.be.device
        # every per-step output goes into buffers allocated BEFORE the loop (and the per-step temporaries are the solver's
        # persistent buffers): an allocation inside the loop that the caching allocator cannot serve from its pool is a
        # cudaMalloc, i.e. a device synchronisation — measured at 10-14 ms per closed-loop step of 131072 scenarios
        traj = torch.empty((steps + 1, B, nx), device=dev, dtype=self.dtype) if record else None
        us = torch.empty((steps, B, nu), device=dev, dtype=self.dtype)
        its = torch.empty((steps, B), device=dev, dtype=torch.int32)
        x_pp = [s.buffer("cl_x0", tuple(x_em.shape)), s.buffer("cl_x1", tuple(x_em.shape))]
        if record:
            self.be.check(self.be.lib.mpcb_to_batch_major(_dt(self.dtype), B, nx, s.ld, ptr(x_em), ptr(traj[0]), self.be.stream()))
        dt = _dt(self.dtype)
        for k in range(steps):
            if k > 0:
                if bounds_at is not None and bounds_at(k):
                    self.update_bounds(**bounds_at(k))
                s.update(x_init=x_em, element_major=True)
                s.solve()
                _, _, u = s.solution(want_x=False, want_y=False, want_u=True, reuse=True)
                info = s.info(reuse=True)
            else:
                u, info = res.u, res.info
            us[k].copy_(u[:, 0, :]); its[k].copy_(info.iter)
            x_next = x_pp[k & 1]
            self.be.check(self.be.lib.mpcb_plant_step(dt, B, s.ld, nx, nu, int(self.shared), ptr(A), ptr(Bm), ptr(None),
                                                      ptr(x_em), ptr(u), N * nu, ptr(x_next), self.be.stream()))
            x_em = x_next
            if record:
                self.be.check(self.be.lib.mpcb_to_batch_major(dt, B, nx, s.ld, ptr(x_em), ptr(traj[k + 1]), self.be.stream()))
        return traj, us, its

    def solve(self, state, reference, speed=None):
        """Single vehicle: returns the input sequence (N, nu) as numpy.  Raises like the reference
        (vehicle_lateral_mpc_slack_increment.py:252-253) if OSQP's status is not 'solved'.
        The first call sets the problem up; later calls are prob.update() + warm-started prob.solve()."""
        st = np.asarray(state, dtype=np.float64).reshape(1, self.nx)
        ref = np.asarray(reference, dtype=np.float64).reshape(1, -1)
        if not self._is_setup or not self.shared and speed is not None and speed != getattr(self, "_speed", None):
            res = self.solve_batch(st, ref, None if speed is None else np.array([speed], dtype=np.float64))
            self._speed = speed
        else:
            res = self.update_batch(st, ref)
        self.last = res
        if int(res.info.status_val[0]) != 1:
            raise ValueError("OSQP did not solve the problem!")
        return res.u[0].cpu().numpy().astype(np.float64)
