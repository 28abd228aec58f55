"""ORACLE — test infrastructure only.  Never imported by the product path.

CPU restatement of the reference's "cast MPC problem to a QP" step, for every
formulation the reference has.  All of them are instances of one stage-structured
QP (the *canonical MPC-QP*), which is also exactly what the CUDA build kernel
(build_one in python-mpc_b200/csrc/shape_ops.cuh, behind mpcb_build_qp) emits:

  variables  v = [x_0 .. x_N | u_0 .. u_{N-1} | s_0 .. s_N]        (s only if slack)
  cost       sum_k 1/2 x_k' Q_k x_k - (Q_k xr_k)' x_k + 1/2 u_k' R u_k + 1/2 s_k' W s_k
             (Q_k = Q for k < N, QN for k = N; all weights diagonal)
  rows       dyn_0      : -x_0                              = -x_init
             dyn_{k+1}  : A_k x_k + B_k u_k - x_{k+1}       = -g_k
             bx_k       : xmin_k <= x_k + S s_k <= xmax_k
             bu_k       : umin   <= u_k         <= umax

Reference call sites restated:
  vanilla, time-invariant  Control/MPC/mpc_kinematics.py:150-213          (mpc)
  vanilla, stage boxes     Control/MPC/mpc_kinematics.py:215-270          (mpc_)
  vanilla, time-varying    Control/MPC/mpc_dynamics.py:156-252            (mpc)
                           Control/MPC/mpc_kinematics_pred_matrix.py:268-353 (mpc__)
  incremental (delta-u)    Control/MPC/mpc_dynamics.py:284-402            (mpc_increment)
                           Control/MPC/mpc_incre_kine_func.py:84-186
  slack + incremental      vehicle_lateral_mpc_slack_increment.py:32-121, 132-229

The reference builds these with scipy.sparse kron/hstack/vstack; this file builds
the same matrices from explicit (row, col, val) triplets so that it is an
independent statement of the layout.  tests/test_oracle.py checks it against
the matrices captured from the reference's own code (tests/golden/).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class MpcQp:
    """Canonical stage-structured MPC-QP (host-side, one problem)."""
    N: int
    nx: int
    nu: int
    A: np.ndarray        # (N, nx, nx)
    B: np.ndarray        # (N, nx, nu)
    g: np.ndarray        # (N, nx)
    Q: np.ndarray        # (nx,)  diagonal stage weight
    QN: np.ndarray       # (nx,)  diagonal terminal weight
    R: np.ndarray        # (nu,)
    Xr: np.ndarray       # (nx, N+1) stage references
    xmin: np.ndarray     # (N+1, nx)
    xmax: np.ndarray     # (N+1, nx)
    umin: np.ndarray     # (nu,)
    umax: np.ndarray     # (nu,)
    x_init: np.ndarray   # (nx,)
    slack: bool = False
    W: np.ndarray = field(default=None)   # (nx,) slack cost (W and WN are equal in the reference)
    S: np.ndarray = field(default=None)   # (nx,) slack coupling (weight_slack_tilda)

    @property
    def ns(self):
        return self.nx if self.slack else 0

    @property
    def nvar(self):
        return (self.N + 1) * self.nx + self.N * self.nu + (self.N + 1) * self.ns

    @property
    def ncon(self):
        return 2 * (self.N + 1) * self.nx + self.N * self.nu

    # index helpers (reference ordering)
    def ix(self, k, i=0): return k * self.nx + i
    def iu(self, k, i=0): return (self.N + 1) * self.nx + k * self.nu + i
    def is_(self, k, i=0): return (self.N + 1) * self.nx + self.N * self.nu + k * self.nx + i
    def rdyn(self, k, i=0): return k * self.nx + i
    def rbx(self, k, i=0): return (self.N + 1) * self.nx + k * self.nx + i
    def rbu(self, k, i=0): return 2 * (self.N + 1) * self.nx + k * self.nu + i


def _bcast_stage(M, N):
    M = np.asarray(M, dtype=np.float64)
    return np.broadcast_to(M, (N,) + M.shape[-2:]).copy() if M.ndim == 2 else M.copy()


def canonical(N, A, B, g, Q, QN, R, Xr, xmin, xmax, umin, umax, x_init,
              slack=False, W=None, S=None) -> MpcQp:
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if A.ndim == 2:
        A = np.broadcast_to(A, (N,) + A.shape).copy()
    if B.ndim == 2:
        B = np.broadcast_to(B, (N,) + B.shape).copy()
    nx, nu = B.shape[1], B.shape[2]
    g = np.zeros((N, nx)) if g is None else np.asarray(g, dtype=np.float64).reshape(-1, nx)
    if g.shape[0] == 1:
        g = np.broadcast_to(g, (N, nx)).copy()
    Xr = np.asarray(Xr, dtype=np.float64)
    if Xr.ndim == 1:
        Xr = np.tile(Xr[:, None], (1, N + 1))
    xmin = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (N + 1, nx)).copy()
    xmax = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (N + 1, nx)).copy()
    d = lambda v: np.asarray(v.diagonal() if hasattr(v, "diagonal") and getattr(v, "ndim", 1) == 2 else v,
                             dtype=np.float64).ravel()
    return MpcQp(N=N, nx=nx, nu=nu, A=A, B=B, g=g, Q=d(Q), QN=d(QN), R=d(R), Xr=Xr,
                 xmin=xmin, xmax=xmax, umin=np.asarray(umin, dtype=np.float64).ravel(),
                 umax=np.asarray(umax, dtype=np.float64).ravel(),
                 x_init=np.asarray(x_init, dtype=np.float64).ravel(), slack=slack,
                 W=None if W is None else d(W), S=None if S is None else d(S))


def augment_increment(Ad, Bd, gd):
    """delta-u augmentation: x~ = [x; u_prev], input delta-u.
    A~ = [[Ad, Bd], [0, I]], B~ = [[Bd], [I]], g~ = [gd; 0]
    (mpc_dynamics.py:337-341, 388-391; vehicle_lateral_mpc_slack_increment.py:48-52)."""
    Ad = np.asarray(Ad, dtype=np.float64); Bd = np.asarray(Bd, dtype=np.float64)
    lead = Ad.shape[:-2]
    nx, nu = Bd.shape[-2:]
    At = np.zeros(lead + (nx + nu, nx + nu)); Bt = np.zeros(lead + (nx + nu, nu))
    At[..., :nx, :nx] = Ad; At[..., :nx, nx:] = Bd; At[..., nx:, nx:] = np.eye(nu)
    Bt[..., :nx, :] = Bd; Bt[..., nx:, :] = np.eye(nu)
    gt = None
    if gd is not None:
        gd = np.asarray(gd, dtype=np.float64).reshape(lead + (nx,))
        gt = np.zeros(lead + (nx + nu,)); gt[..., :nx] = gd
    return At, Bt, gt


def assemble(p: MpcQp):
    """(P, q, A, l, u) of the canonical MPC-QP in the reference's ordering."""
    N, nx, nu, ns = p.N, p.nx, p.nu, p.ns
    n, m = p.nvar, p.ncon
    Pd = np.zeros(n); q = np.zeros(n)
    for k in range(N + 1):
        w = p.QN if k == N else p.Q
        Pd[p.ix(k):p.ix(k) + nx] = w
        q[p.ix(k):p.ix(k) + nx] = -w * p.Xr[:, k]
    for k in range(N):
        Pd[p.iu(k):p.iu(k) + nu] = p.R
    if ns:
        for k in range(N + 1):
            Pd[p.is_(k):p.is_(k) + nx] = p.W
    rows, cols, vals = [], [], []
    l = np.zeros(m); u = np.zeros(m)
    for k in range(N + 1):
        for i in range(nx):
            rows.append(p.rdyn(k, i)); cols.append(p.ix(k, i)); vals.append(-1.0)
        if k == 0:
            l[p.rdyn(0):p.rdyn(0) + nx] = -p.x_init
        else:
            for i in range(nx):
                for j in range(nx):
                    rows.append(p.rdyn(k, i)); cols.append(p.ix(k - 1, j)); vals.append(p.A[k - 1, i, j])
                for j in range(nu):
                    rows.append(p.rdyn(k, i)); cols.append(p.iu(k - 1, j)); vals.append(p.B[k - 1, i, j])
            l[p.rdyn(k):p.rdyn(k) + nx] = -p.g[k - 1]
    u[:(N + 1) * nx] = l[:(N + 1) * nx]
    for k in range(N + 1):
        for i in range(nx):
            rows.append(p.rbx(k, i)); cols.append(p.ix(k, i)); vals.append(1.0)
            if ns:
                rows.append(p.rbx(k, i)); cols.append(p.is_(k, i)); vals.append(p.S[i])
        l[p.rbx(k):p.rbx(k) + nx] = p.xmin[k]
        u[p.rbx(k):p.rbx(k) + nx] = p.xmax[k]
    for k in range(N):
        for i in range(nu):
            rows.append(p.rbu(k, i)); cols.append(p.iu(k, i)); vals.append(1.0)
        l[p.rbu(k):p.rbu(k) + nu] = p.umin
        u[p.rbu(k):p.rbu(k) + nu] = p.umax
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsc()
    P = sp.diags(Pd).tocsc()
    return P, q, A, l, u


# ---------------------------------------------------------------- reference-signature helpers
def qp_vanilla(Ad, Bd, gd, x_vec, Xr, Q, QN, R, N, xmin, xmax, umin, umax):
    """mpc(...) of mpc_kinematics.py:150 (matrices) and mpc_dynamics.py:156 (lists)."""
    if isinstance(Ad, (list, tuple)):
        Ad = np.stack([np.asarray(a) for a in Ad]); Bd = np.stack([np.asarray(b) for b in Bd])
        gd = np.stack([np.asarray(g).reshape(-1) for g in gd])
    else:
        gd = None if gd is None else np.asarray(gd).reshape(1, -1)
    return canonical(N, Ad, Bd, gd, Q, QN, R, Xr, xmin, xmax, umin, umax, x_vec)


def qp_increment(Ad_list, Bd_list, gd_list, x_tilda_vec, Xr, Q, QN, R, N,
                 xmin_tilda, xmax_tilda, del_umin, del_umax):
    """mpc_increment(...) of mpc_dynamics.py:284 / mpc_incre_kine_func.py:84."""
    Ad = np.stack([np.asarray(a) for a in Ad_list]); Bd = np.stack([np.asarray(b) for b in Bd_list])
    gd = np.stack([np.asarray(g).reshape(-1) for g in gd_list])
    nx, nu = Bd.shape[1:]
    At, Bt, gt = augment_increment(Ad, Bd, gd)
    z = np.zeros(nu)
    d = lambda v: np.asarray(v.diagonal() if getattr(v, "ndim", 1) == 2 else v, dtype=np.float64).ravel()
    Xr = np.asarray(Xr, dtype=np.float64)
    Xrt = np.vstack([Xr, np.zeros((nu, Xr.shape[1]))])
    return canonical(N, At, Bt, gt, np.concatenate([d(Q), z]), np.concatenate([d(QN), z]), R, Xrt,
                     xmin_tilda, xmax_tilda, del_umin, del_umax, x_tilda_vec)


def qp_slack_increment(Ad_sys, Bd_sys, x0_tilda, xr, Q, R, W_tilda, weight_slack_tilda, N,
                       xmin_tilda, xmax_tilda, del_umin, del_umax):
    """vehicle_lateral_mpc_slack_increment.py:32-115 (QN = Q~, WN = W~)."""
    At, Bt, _ = augment_increment(Ad_sys, Bd_sys, None)
    nu = Bt.shape[1]
    d = lambda v: np.asarray(v.diagonal() if getattr(v, "ndim", 1) == 2 else v, dtype=np.float64).ravel()
    Qt = np.concatenate([d(Q), np.zeros(nu)])
    xrt = np.concatenate([np.asarray(xr, dtype=np.float64).ravel(), np.zeros(nu)])
    return canonical(N, At, Bt, None, Qt, Qt, R, xrt, xmin_tilda, xmax_tilda, del_umin, del_umax,
                     x0_tilda, slack=True, W=W_tilda, S=weight_slack_tilda)
