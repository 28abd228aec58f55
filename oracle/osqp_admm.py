"""ORACLE — test infrastructure only.  Never imported by the product path.

Restatement, in numpy/scipy, of the OSQP ADMM algorithm that the reference
calls through ``osqp.OSQP().setup(...); .solve()``:

    /root/reference/Control/MPC/mpc_kinematics.py:205-211
    /root/reference/Control/MPC/mpc_dynamics.py:248-252, 398-402
    /root/reference/vehicle_lateral_mpc_slack_increment.py:118-121, 237-248

OSQP itself is a third-party dependency that is NOT vendored in /root/reference
and is NOT installable in this image (no network, not in /opt/wheelhouse).  The
reference does not pin a version; its files are dated 2019-08/09, which is the
osqp 0.6.x series (0.6.1, QDLDL backend).  This file therefore restates the
PUBLISHED algorithm:

    B. Stellato, G. Banjac, P. Goulart, A. Bemporad, S. Boyd,
    "OSQP: an operator splitting solver for quadratic programs",
    Math. Prog. Comp. 12 (2020) — Algorithm 1 (ADMM), Algorithm 2 (modified
    Ruiz equilibration), section 3.4 (termination), section 5.2 (rho per
    constraint type),

with the constants and the evaluation order of the 0.6.x C sources
(osqp/src/{osqp.c,auxil.c,scaling.c}, osqp/include/constants.h).

PARITY UNPINNED against an OSQP binary: no OSQP build exists in this image, so
no golden iterates from the real solver could be generated.  What the oracle IS
pinned to (tests/test_oracle.py):
  * the QP data (P, q, A, l, u) produced by the reference's own assembly code,
    executed here from /root/reference (tests/golden/*.npz, oracle/make_golden.py);
  * OSQP's documented demo QP (docs "Setup and solve" example: x* = [0.3, 0.7]);
  * KKT optimality of the converged point, checked independently of ADMM;
  * agreement between this file and the independent C port oracle/osqp_admm.c.

The defaults below are OSQP 0.6.x defaults; BASELINE.json's north_star asks for
adaptive_rho and polish off, which is how every parity test calls it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# osqp/include/constants.h (0.6.x)
OSQP_INFTY = 1e30
RHO_MIN = 1e-6
RHO_MAX = 1e6
RHO_EQ_OVER_RHO_INEQ = 1e3
RHO_TOL = 1e-4
MIN_SCALING = 1e-4
MAX_SCALING = 1e4

OSQP_SOLVED = 1
OSQP_SOLVED_INACCURATE = 2
OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3
OSQP_DUAL_INFEASIBLE_INACCURATE = 4
OSQP_MAX_ITER_REACHED = -2
OSQP_PRIMAL_INFEASIBLE = -3
OSQP_DUAL_INFEASIBLE = -4
OSQP_UNSOLVED = -10

STATUS_STRING = {
    OSQP_SOLVED: "solved",
    OSQP_SOLVED_INACCURATE: "solved inaccurate",
    OSQP_PRIMAL_INFEASIBLE: "primal infeasible",
    OSQP_PRIMAL_INFEASIBLE_INACCURATE: "primal infeasible inaccurate",
    OSQP_DUAL_INFEASIBLE: "dual infeasible",
    OSQP_DUAL_INFEASIBLE_INACCURATE: "dual infeasible inaccurate",
    OSQP_MAX_ITER_REACHED: "maximum iterations reached",
    OSQP_UNSOLVED: "unsolved",
}

DEFAULT_SETTINGS = dict(
    rho=0.1, sigma=1e-6, alpha=1.6, max_iter=4000,
    eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4, eps_dual_inf=1e-4,
    scaling=10, check_termination=25, warm_start=True,
    adaptive_rho=False, polish=False, scaled_termination=False,
)


def _limit_scaling(v):
    """scaling.c: limit_scaling — zero-norm -> 1, cap at MAX_SCALING."""
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


class Info:
    pass


class Results:
    pass


class OSQP:
    """Mirror of the ``osqp.OSQP`` object surface the reference uses:
    setup / update(q=, l=, u=) / warm_start / solve -> results.x, .y,
    .info.status, .info.status_val, .info.iter."""

    def setup(self, P, q, A, l, u, **settings):
        s = dict(DEFAULT_SETTINGS)
        for k, v in settings.items():
            if k in ("verbose", "linsys_solver", "time_limit", "delta",
                     "polish_refine_iter", "adaptive_rho_interval",
                     "adaptive_rho_tolerance", "adaptive_rho_fraction"):
                continue
            if k not in s:
                raise TypeError("unknown OSQP setting %r" % k)
            s[k] = v
        if s["adaptive_rho"] or s["polish"]:
            raise NotImplementedError(
                "oracle covers the north_star configuration: adaptive_rho=False, polish=False")
        self.s = s
        P = sp.csc_matrix(P, dtype=np.float64)
        A = sp.csc_matrix(A, dtype=np.float64)
        self.n = P.shape[0]
        self.m = A.shape[0]
        # python interface: only the upper triangle of P is used, infinities clipped
        self.P = (sp.triu(P, k=0) + sp.triu(P, k=1).T).tocsc()
        self.A = A
        self.q = np.array(q, dtype=np.float64).copy()
        self.l = np.maximum(np.array(l, dtype=np.float64), -OSQP_INFTY)
        self.u = np.minimum(np.array(u, dtype=np.float64), OSQP_INFTY)
        s["rho"] = min(max(s["rho"], RHO_MIN), RHO_MAX)
        self._scale()
        self._set_rho_vec()
        self._factor()
        # cold start (osqp.c: cold_start)
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        return self

    # ---------------------------------------------------------------- scaling
    def _scale(self):
        """scaling.c: scale_data (Algorithm 2 of the paper)."""
        n, m = self.n, self.m
        P, A, q = self.P.copy(), self.A.copy(), self.q.copy()
        D = np.ones(n)
        E = np.ones(m)
        c = 1.0
        for _ in range(self.s["scaling"]):
            absP, absA = abs(P), abs(A)
            Pn = absP.max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
            An = absA.max(axis=0).toarray().ravel() if A.nnz else np.zeros(n)
            Dt = np.maximum(Pn, An)
            Et = absA.max(axis=1).toarray().ravel() if A.nnz else np.zeros(m)
            Dt = 1.0 / np.sqrt(_limit_scaling(Dt))
            Et = 1.0 / np.sqrt(_limit_scaling(Et))
            P = sp.diags(Dt) @ P @ sp.diags(Dt)
            A = sp.diags(Et) @ A @ sp.diags(Dt)
            q = Dt * q
            D *= Dt
            E *= Et
            # cost normalisation
            Pn = abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
            c_temp = Pn.mean()
            inf_norm_q = float(_limit_scaling(np.array([np.abs(q).max()]))[0])
            c_temp = max(c_temp, inf_norm_q)
            c_temp = float(_limit_scaling(np.array([c_temp]))[0])
            c_temp = 1.0 / c_temp
            P = P * c_temp
            q = q * c_temp
            c *= c_temp
        self.D, self.E, self.c = D, E, c
        self.Dinv, self.Einv, self.cinv = 1.0 / D, 1.0 / E, 1.0 / c
        self.Ps, self.As, self.qs = sp.csc_matrix(P), sp.csc_matrix(A), q
        self.ls = E * self.l
        self.us = E * self.u

    def _set_rho_vec(self):
        """auxil.c: set_rho_vec — evaluated on the SCALED bounds."""
        rho = self.s["rho"]
        l, u = self.ls, self.us
        unc = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
        eq = (~unc) & (u - l < RHO_TOL)
        self.constr_type = np.where(unc, -1, np.where(eq, 1, 0))
        self.rho_vec = np.where(unc, RHO_MIN, np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho, rho))
        self.rho_inv_vec = 1.0 / self.rho_vec

    def _factor(self):
        """KKT = [[P + sigma I, A'], [A, -diag(1/rho)]] (paper eq. (17)).
        OSQP factors it with QDLDL; any exact solve gives the same iterates."""
        n = self.n
        K = sp.bmat([[self.Ps + self.s["sigma"] * sp.eye(n), self.As.T],
                     [self.As, -sp.diags(self.rho_inv_vec)]], format="csc")
        self._lu = spla.splu(K)

    # ---------------------------------------------------------------- updates
    def update(self, q=None, l=None, u=None):
        """osqp.c: osqp_update_lin_cost / osqp_update_bounds — new data are scaled
        with the EXISTING D, E, c; scaling is not recomputed."""
        if q is not None:
            self.q = np.array(q, dtype=np.float64).copy()
            self.qs = self.c * self.D * self.q
        if l is not None:
            self.l = np.maximum(np.array(l, dtype=np.float64), -OSQP_INFTY)
            self.ls = self.E * self.l
        if u is not None:
            self.u = np.minimum(np.array(u, dtype=np.float64), OSQP_INFTY)
            self.us = self.E * self.u
        if l is not None or u is not None:
            if np.any(self.ls > self.us):
                raise ValueError("lower bound must be lower than or equal to upper bound")
            # osqp_update_bounds: update_rho_vec refactors when a constraint type changed
            old = self.constr_type.copy()
            self._set_rho_vec()
            if np.any(old != self.constr_type):
                self._factor()

    def warm_start(self, x=None, y=None):
        """osqp.c: osqp_warm_start — scale x, y; z = A x."""
        if x is not None:
            self.x = self.Dinv * np.asarray(x, dtype=np.float64)
        if y is not None:
            self.y = self.Einv * np.asarray(y, dtype=np.float64) * self.c
        self.z = self.As @ self.x

    # ---------------------------------------------------------------- solve
    def _residuals(self):
        x, z, y = self.x, self.z, self.y
        Ax = self.As @ x
        Px = self.Ps @ x
        Aty = self.As.T @ y
        pri_res = np.abs(self.Einv * (Ax - z)).max() if self.m else 0.0
        dua_res = self.cinv * np.abs(self.Dinv * (self.qs + Aty + Px)).max()
        return pri_res, dua_res, Ax, Px, Aty

    def _check_termination(self, approximate, dx, dy):
        s = self.s
        eps_abs, eps_rel = s["eps_abs"], s["eps_rel"]
        eps_pinf, eps_dinf = s["eps_prim_inf"], s["eps_dual_inf"]
        if approximate:
            eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10
        pri_res, dua_res, Ax, Px, Aty = self._residuals()
        self.pri_res, self.dua_res = pri_res, dua_res
        prim_ok = True
        if self.m:
            eps_prim = eps_abs + eps_rel * max(np.abs(self.Einv * self.z).max(),
                                               np.abs(self.Einv * Ax).max())
            prim_ok = pri_res < eps_prim
            if not prim_ok and self._is_primal_infeasible(eps_pinf, dy):
                return OSQP_PRIMAL_INFEASIBLE_INACCURATE if approximate else OSQP_PRIMAL_INFEASIBLE
        eps_dual = eps_abs + eps_rel * self.cinv * max(
            np.abs(self.Dinv * self.qs).max(), np.abs(self.Dinv * Aty).max(),
            np.abs(self.Dinv * Px).max())
        dual_ok = dua_res < eps_dual
        if not dual_ok and self._is_dual_infeasible(eps_dinf, dx):
            return OSQP_DUAL_INFEASIBLE_INACCURATE if approximate else OSQP_DUAL_INFEASIBLE
        if prim_ok and dual_ok:
            return OSQP_SOLVED_INACCURATE if approximate else OSQP_SOLVED
        return None

    def _is_primal_infeasible(self, eps, dy):
        """auxil.c: is_primal_infeasible (paper eq. (22))."""
        dy = dy.copy()
        up_inf = self.us > OSQP_INFTY * MIN_SCALING
        lo_inf = self.ls < -OSQP_INFTY * MIN_SCALING
        dy[up_inf & lo_inf] = 0.0
        m1 = up_inf & ~lo_inf
        dy[m1] = np.minimum(dy[m1], 0.0)
        m2 = lo_inf & ~up_inf
        dy[m2] = np.maximum(dy[m2], 0.0)
        norm_dy = np.abs(self.E * dy).max()
        if norm_dy > eps:
            lhs = np.sum(self.us * np.maximum(dy, 0.0) + self.ls * np.minimum(dy, 0.0))
            if lhs < -eps * norm_dy:
                Atdy = self.Dinv * (self.As.T @ dy)
                return np.abs(Atdy).max() < eps * norm_dy
        return False

    def _is_dual_infeasible(self, eps, dx):
        """auxil.c: is_dual_infeasible (paper eq. (24))."""
        norm_dx = np.abs(self.D * dx).max()
        cost_scaling = self.c
        if norm_dx > eps:
            if self.qs @ dx < -cost_scaling * eps * norm_dx:
                Pdx = self.Dinv * (self.Ps @ dx)
                if np.abs(Pdx).max() < cost_scaling * eps * norm_dx:
                    Adx = self.Einv * (self.As @ dx)
                    for i in range(self.m):
                        if ((self.us[i] < OSQP_INFTY * MIN_SCALING and Adx[i] > eps * norm_dx) or
                                (self.ls[i] > -OSQP_INFTY * MIN_SCALING and Adx[i] < -eps * norm_dx)):
                            return False
                    return True
        return False

    def iterate(self):
        """One ADMM iteration — osqp.c: update_xz_tilde, update_x, update_z, update_y
        (paper Algorithm 1, steps 3-7)."""
        s = self.s
        n = self.n
        alpha, sigma = s["alpha"], s["sigma"]
        x_prev, z_prev = self.x, self.z
        rhs = np.concatenate([sigma * x_prev - self.qs,
                              z_prev - self.rho_inv_vec * self.y])
        sol = self._lu.solve(rhs)
        xt = sol[:n]
        zt = z_prev + self.rho_inv_vec * (sol[n:] - self.y)
        x = alpha * xt + (1.0 - alpha) * x_prev
        zr = alpha * zt + (1.0 - alpha) * z_prev
        z = np.minimum(np.maximum(zr + self.rho_inv_vec * self.y, self.ls), self.us)
        dy = self.rho_vec * (zr - z)
        self.y = self.y + dy
        dx = x - x_prev
        self.x, self.z = x, z
        return dx, dy

    def solve(self, trace=None):
        s = self.s
        if not s["warm_start"]:
            self.x = np.zeros(self.n); self.z = np.zeros(self.m); self.y = np.zeros(self.m)
        status = OSQP_UNSOLVED
        it = 0
        dx = np.zeros(self.n); dy = np.zeros(self.m)
        checked = False
        for it in range(1, s["max_iter"] + 1):
            dx, dy = self.iterate()
            if trace is not None:
                trace.append((self.D * self.x, self.Einv * self.z, self.cinv * self.E * self.y))
            checked = False
            if s["check_termination"] and it % s["check_termination"] == 0:
                checked = True
                st = self._check_termination(False, dx, dy)
                if st is not None:
                    status = st
                    break
        if status == OSQP_UNSOLVED:
            if not checked:
                st = self._check_termination(False, dx, dy)
                if st is not None:
                    status = st
            if status == OSQP_UNSOLVED:
                st = self._check_termination(True, dx, dy)
                status = st if st is not None else OSQP_MAX_ITER_REACHED
        res = Results()
        res.x = self.D * self.x
        res.y = self.cinv * self.E * self.y
        res.z = self.Einv * self.z
        res.info = Info()
        res.info.iter = it
        res.info.status_val = status
        res.info.status = STATUS_STRING[status]
        res.info.pri_res = getattr(self, "pri_res", np.nan)
        res.info.dua_res = getattr(self, "dua_res", np.nan)
        xr = res.x
        res.info.obj_val = 0.5 * xr @ (self.P @ xr) + self.q @ xr
        if status in (OSQP_PRIMAL_INFEASIBLE, OSQP_PRIMAL_INFEASIBLE_INACCURATE,
                      OSQP_DUAL_INFEASIBLE, OSQP_DUAL_INFEASIBLE_INACCURATE):
            res.x = np.full(self.n, np.nan)
            res.y = np.full(self.m, np.nan)
        return res


def solve_qp(P, q, A, l, u, **settings):
    return OSQP().setup(P, q, A, l, u, **settings).solve()
