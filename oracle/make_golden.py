"""ORACLE — test infrastructure only.  Generates tests/golden/*.npz.

Runs the UNMODIFIED reference code from /root/reference in this container (it is
not present on the GPU box, so the outputs are committed as fixtures):

  * Vehicle_Dynamics/vehicle_models.py is imported as-is (matplotlib, which is not
    installed, is replaced by an inert stub — it is only used for plotting);
  * Control/MPC/{mpc_kinematics,mpc_dynamics,mpc_incre_kine_func,mpc_kinematics_pred_matrix,
    mpc_increment_kinematics_pred_matrix}.py are imported as-is and their mpc() / mpc_() / mpc__() /
    mpc_increment() functions are called on seeded inputs;
  * vehicle_lateral_mpc_slack_increment.py is a top-level script; it is executed from its source text twice:
      - with two literals changed (N = 100 -> 20, the horizon BASELINE.json names, and nsim = 1500 -> 120):
        the H = 20 QP data of configs[0] (P, q, A, l, u bit-exact) and the head of its closed loop.  The H = 20 loop
        cannot be run further: with the script's own weights and rate limits a 20-step horizon does not stabilise
        the plant (|e_y| grows past 30 m by step 360) and a solve returns "solved inaccurate" at step 146..374
        whatever the fixed rho, i.e. the script itself raises;
      - UNMODIFIED (N = 100, nsim = 1500): the whole closed loop of the script, including the bound switches of
        lines 158-172 (xmin_tilda[3] = 2 for steps 401..900, applied by prob.update(q, l, u) at line 237) and the
        slack-active regime they cause (e_y settles at 1.0, slack ~ 1).  Solver settings: the script's own
        (OSQP defaults, eps 1e-3) with adaptive_rho/polish off as north_star fixes them and rho = 10 — with
        adaptation off the default rho = 0.1 does not converge within max_iter at this horizon.

`import osqp` inside the reference resolves to a recording shim: it stores the
(P, q, A, l, u) the reference assembled — these are REAL reference outputs — and
solves with oracle/osqp_admm.py (adaptive_rho and polish off), so the solver
outputs stored next to them are "reference assembly + oracle ADMM".

    python oracle/make_golden.py          # rewrites tests/golden/
"""
import io
import os
import sys
import types
import contextlib

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import osqp_admm  # noqa: E402


# ----------------------------------------------------------------------------- shims
class _Inert:
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


def _install_shims(records, solver_settings):
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: _Inert()
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt

    class RecordingOSQP:
        def setup(self, P, q, A, l, u, **kw):
            rec = dict(P=sp.csc_matrix(P).toarray(), q=np.array(q, dtype=float),
                       A=sp.csc_matrix(A).toarray(), l=np.array(l, dtype=float),
                       u=np.array(u, dtype=float), kw=dict(kw), updates=[], results=[])
            records.append(rec)
            self._rec = rec
            kw = {k: v for k, v in kw.items() if k in ("warm_start",)}
            kw.update(solver_settings)
            self._o = osqp_admm.OSQP().setup(P, q, A, l, u, **kw)

        def update(self, **kw):
            self._rec["updates"].append({k: np.array(v, dtype=float) for k, v in kw.items()})
            self._o.update(**kw)

        def solve(self):
            r = self._o.solve()
            self._rec["results"].append(dict(x=r.x.copy(), y=r.y.copy(), iter=r.info.iter,
                                             status_val=r.info.status_val))
            return r

    osqp = types.ModuleType("osqp"); osqp.OSQP = RecordingOSQP
    sys.modules["osqp"] = osqp


def _import_ref(relpath, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    records = []
    settings = dict(adaptive_rho=False, polish=False, eps_abs=1e-4, eps_rel=1e-4)
    _install_shims(records, settings)
    sys.path.insert(0, os.path.join(REF, "Vehicle_Dynamics"))
    vm = _import_ref("Vehicle_Dynamics/vehicle_models.py", "vehicle_models")
    sys.modules["vehicle_models"] = vm
    rng = np.random.default_rng(20191)

    # ---------------------------------------------------------------- vehicle models
    veh = vm.Vehicle_Dynamics(m=1300, l_f=1.25, l_r=1.40, width=1.78, length=4.25, turning_circle=10.4,
                              C_d=0.34, A_f=2.0, C_roll=0.015, dt=0.05)
    K = 64
    xs = np.zeros((K, 6)); us = np.zeros((K, 2))
    xs[:, 0:2] = rng.uniform(-50, 50, (K, 2))
    xs[:, 2] = rng.uniform(-3.0, 3.0, K)
    xs[:, 3] = rng.uniform(1.0, 30.0, K)
    xs[:, 4] = rng.uniform(-1.0, 1.0, K)
    xs[:, 5] = rng.uniform(-0.5, 0.5, K)
    us[:, 0] = rng.uniform(-0.25, 0.25, K)
    us[:, 1] = rng.uniform(-3.0, 1.0, K)
    # exercise the low-speed guards of get_dynamics_model (vehicle_models.py:143-159)
    xs[0, 3] = 0.1; xs[1, 3] = 0.4; xs[2, 3] = -0.1; xs[3, 3] = -0.4; xs[4, 3] = -5.0; xs[5, 3] = 0.0
    Ad = np.zeros((K, 6, 6)); Bd = np.zeros((K, 6, 2)); gd = np.zeros((K, 6)); xn = np.zeros((K, 6))
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(K):
            a, b, g = veh.get_dynamics_model(xs[i].copy(), us[i].copy())
            Ad[i], Bd[i], gd[i] = a, b, g[:, 0]
            xn[i] = veh.update_dynamics_model(xs[i].copy(), us[i].copy())[0][:, 0]
    kin = vm.Vehicle_Kinematics(l_f=1.25, l_r=1.40, dt=0.02)
    xk = np.zeros((K, 4)); uk = np.zeros((K, 2))
    xk[:, 0:2] = rng.uniform(-50, 50, (K, 2)); xk[:, 2] = rng.uniform(0.0, 30.0, K)
    xk[:, 3] = rng.uniform(-3.0, 3.0, K)
    uk[:, 0] = rng.uniform(-0.25, 0.25, K); uk[:, 1] = rng.uniform(-3.0, 1.0, K)
    Ak = np.zeros((K, 4, 4)); Bk = np.zeros((K, 4, 2)); Ck = np.zeros((K, 4)); xkn = np.zeros((K, 4))
    for i in range(K):
        a, b, c = kin.get_kinematics_model(xk[i].copy(), uk[i].copy())
        Ak[i], Bk[i], Ck[i] = a, b, c[:, 0]
        xkn[i] = kin.update_kinematics_model(xk[i].copy(), uk[i].copy())
    np.savez_compressed(os.path.join(OUT, "vehicle_models.npz"),
                        dyn_x=xs, dyn_u=us, dyn_Ad=Ad, dyn_Bd=Bd, dyn_gd=gd, dyn_xnext=xn, dyn_dt=0.05,
                        kin_x=xk, kin_u=uk, kin_A=Ak, kin_B=Bk, kin_C=Ck, kin_xnext=xkn, kin_dt=0.02)

    # ---------------------------------------------------------------- mpc_kinematics.mpc (vanilla, LTI)
    sys.path.insert(0, os.path.join(REF, "Control", "MPC"))
    mk = _import_ref("Control/MPC/mpc_kinematics.py", "ref_mpc_kinematics")
    N = 20
    Q = sp.diags([1.0, 1.0, 5.0, 10.0]); QN = sp.diags([10.0, 10.0, 50.0, 50.0]); R = sp.diags([0.1, 0.1])
    umin = np.array([-np.deg2rad(15), -3.]); umax = np.array([np.deg2rad(15), 1.])
    xmin = np.array([-np.inf, -np.inf, -100., -np.pi]); xmax = np.array([np.inf, np.inf, 100., np.pi])
    x = np.array([[0.0], [0.0], [5.0], [np.deg2rad(30)]]); u = np.array([[0.0], [0.01]])
    A_, B_, C_ = kin.get_kinematics_model(x, u)
    Xr = np.zeros((4, N + 1)); Xr[0] = np.linspace(0, 4, N + 1); Xr[1] = -5.0; Xr[2] = 10.0
    records.clear()
    res = mk.mpc(A_, B_, C_, x[:, 0], Xr, Q, QN, R, N, xmin, xmax, umin, umax)
    r = records[0]
    np.savez_compressed(os.path.join(OUT, "qp_vanilla_kinematic.npz"), N=N, Ad=A_, Bd=B_, gd=C_[:, 0],
                        x_init=x[:, 0], Xr=Xr, Q=Q.diagonal(), QN=QN.diagonal(), R=R.diagonal(),
                        xmin=xmin, xmax=xmax, umin=umin, umax=umax,
                        P=r["P"], q=r["q"], A=r["A"], l=r["l"], u=r["u"],
                        sol_x=res.x, sol_y=res.y, sol_iter=res.info.iter, sol_status=res.info.status_val)

    # ---------------------------------------------------------------- mpc_dynamics.mpc / mpc_increment (LTV lists)
    md = _import_ref("Control/MPC/mpc_dynamics.py", "ref_mpc_dynamics")
    N = 20
    Q = sp.diags([100.0, 100.0, 100.0, 50.0, 50.0, 50.0]); QN = sp.diags([1000.0, 1000.0, 1000.0, 500.0, 500.0, 500.0])
    R = sp.diags([50, 50])
    del_umin = np.array([-np.deg2rad(2.0), -0.5]); del_umax = np.array([np.deg2rad(2.0), 0.5])
    xmin_t = np.array([-np.inf, -np.inf, -2 * np.pi, -100., -30., -0.5 * np.pi, -np.deg2rad(15), -3.])
    xmax_t = np.array([np.inf, np.inf, 2 * np.pi, 100., 30., 0.5 * np.pi, np.deg2rad(15), 1.])
    x0 = np.array([0.0, 0.0, 0.0, 15.0, 0.0, 0.0]); u0 = np.array([0.0, 0.0])
    pred = np.zeros((8, N + 1)); pred[:6, 0] = x0
    xk_ = np.concatenate([x0, u0])[:, None]
    Al, Bl, gl = [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(N):
            a, b, g = veh.get_dynamics_model(xk_[:6].copy(), xk_[6:].copy())
            Al.append(a); Bl.append(b); gl.append(g)
            xn_ = a @ xk_[:6] + b @ xk_[6:] + g
            xk_ = np.vstack([xn_, xk_[6:]])
            pred[:, i + 1] = xk_[:, 0]
    Xr = np.zeros((6, N + 1)); Xr[0] = np.linspace(0, 15, N + 1); Xr[1] = np.linspace(0, 1.0, N + 1) + 0.5; Xr[3] = 10.0
    records.clear()
    px = np.zeros((8, N + 1)); pdu = np.zeros((2, N + 1))
    with contextlib.redirect_stdout(io.StringIO()):
        px, pdu = md.mpc_increment(Al, Bl, gl, np.concatenate([x0, u0]), Xr, px, pdu, Q, QN, R, N,
                                   xmin_t, xmax_t, del_umin, del_umax)
    r = records[0]
    np.savez_compressed(os.path.join(OUT, "qp_increment_dynamic.npz"), N=N, Ad=np.stack(Al), Bd=np.stack(Bl),
                        gd=np.stack([g[:, 0] for g in gl]), x_init=np.concatenate([x0, u0]), Xr=Xr,
                        Q=Q.diagonal(), QN=QN.diagonal(), R=R.diagonal(), xmin=xmin_t, xmax=xmax_t,
                        umin=del_umin, umax=del_umax, P=r["P"], q=r["q"], A=r["A"], l=r["l"], u=r["u"],
                        pred_x=px, pred_du=pdu, sol_x=r["results"][0]["x"], sol_iter=r["results"][0]["iter"],
                        sol_status=r["results"][0]["status_val"])
    xmin6 = xmin_t[:6]; xmax6 = xmax_t[:6]; umin2 = xmin_t[6:]; umax2 = xmax_t[6:]
    records.clear()
    px = np.zeros((6, N + 1)); pu = np.zeros((2, N + 1))
    with contextlib.redirect_stdout(io.StringIO()):
        px, pu = md.mpc(Al, Bl, gl, x0, Xr, px, pu, Q, QN, R, N, xmin6, xmax6, umin2, umax2)
    r = records[0]
    np.savez_compressed(os.path.join(OUT, "qp_vanilla_dynamic.npz"), N=N, Ad=np.stack(Al), Bd=np.stack(Bl),
                        gd=np.stack([g[:, 0] for g in gl]), x_init=x0, Xr=Xr,
                        Q=Q.diagonal(), QN=QN.diagonal(), R=R.diagonal(), xmin=xmin6, xmax=xmax6,
                        umin=umin2, umax=umax2, P=r["P"], q=r["q"], A=r["A"], l=r["l"], u=r["u"],
                        pred_x=px, pred_u=pu, sol_x=r["results"][0]["x"], sol_iter=r["results"][0]["iter"],
                        sol_status=r["results"][0]["status_val"])

    # ---------------------------------------------------------------- the lateral slack + delta-u closed loop
    src0 = open(os.path.join(REF, "vehicle_lateral_mpc_slack_increment.py")).read()
    assert src0.count("\nN = 100\n") == 1 and src0.count("\nnsim = 1500\n") == 1

    def run_script(src):
        records.clear()
        glb = {"__name__": "ref_lateral_script"}
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(src, "vehicle_lateral_mpc_slack_increment.py", "exec"), glb)
        r = records[0]
        sols = np.stack([s_["x"] for s_ in r["results"]])
        iters = np.array([s_["iter"] for s_ in r["results"]])
        # updates come in pairs per step: (q, l, u) before the solve, (l, u) after it
        q_up = np.stack([up["q"] for up in r["updates"] if "q" in up])
        l_up = np.stack([up["l"] for up in r["updates"] if "q" in up])
        u_up = np.stack([up["u"] for up in r["updates"] if "q" in up])
        return glb, r, sols, iters, q_up, l_up, u_up

    NSIM = 120
    glb, r, sols, iters, q_up, l_up, u_up = run_script(
        src0.replace("\nN = 100\n", "\nN = 20\n").replace("\nnsim = 1500\n", "\nnsim = %d\n" % NSIM))
    np.savez_compressed(os.path.join(OUT, "lateral_slack_increment_closed_loop.npz"), N=20, nsim=NSIM,
                        Ad=glb["Ad_sys"].toarray(), Bd=glb["Bd_sys"].toarray(),
                        P=r["P"], q=r["q"], A=r["A"], l=r["l"], u=r["u"],
                        q_updates=q_up[:4], l_updates=l_up, u_updates=u_up[:4],
                        x1=np.array(glb["plt_x_1"]), x2=np.array(glb["plt_x_2"]), x3=np.array(glb["plt_x_3"]),
                        x4=np.array(glb["plt_x_4"]), u_applied=np.array(glb["plt_u"]),
                        del_u=np.array(glb["plt_del_u"]).ravel(), slack=np.array(glb["plt_s"]),
                        sol_first=sols[0], sol_last=sols[-1], iters=iters)
    iters20 = iters

    # the UNMODIFIED script: N = 100, 1500 steps, bound switches at 401 / 901
    FULL = dict(adaptive_rho=False, polish=False, rho=10.0)        # eps_abs = eps_rel = 1e-3: OSQP's defaults, as the script
    settings_keep = dict(settings)
    settings.clear(); settings.update(FULL)
    glb, r, sols, iters, q_up, l_up, u_up = run_script(src0)
    settings.clear(); settings.update(settings_keep)
    N_full, nsim_full = int(glb["N"]), int(glb["nsim"])
    assert N_full == 100 and nsim_full == 1500 and len(iters) == nsim_full
    # the restated assembly (oracle/ref_qp.py) equals what the script assembled at this horizon too
    from oracle import ref_qp
    DEG = np.pi / 180
    pq = ref_qp.qp_slack_increment(glb["Ad_sys"].toarray(), glb["Bd_sys"].toarray(), np.array([0., 0., 5 * DEG, 3., 0.]),
                                   np.zeros(4), [5., 5., 10., 10.], [10.], [10., 10., 10., 10., 0.], [1., 1., 1., 1., 0.],
                                   N_full, np.array([-np.pi, -0.5 * np.pi, -15 * DEG, -10., -30 * DEG]),
                                   np.array([np.pi, 0.5 * np.pi, 15 * DEG, 10., 30 * DEG]), [-0.5 * DEG], [0.5 * DEG])
    Pq, qq, Aq, lq, uq = ref_qp.assemble(pq)
    assert np.array_equal(Pq.toarray(), r["P"]) and np.array_equal(Aq.toarray(), r["A"]) and np.array_equal(lq, r["l"])
    keep_steps = np.array([0, 1, 400, 401, 402, 900, 901, 902, nsim_full - 1])      # around the bound switches
    np.savez_compressed(os.path.join(OUT, "lateral_slack_increment_closed_loop_full.npz"), N=N_full, nsim=nsim_full,
                        rho=FULL["rho"], eps=1e-3, Ad=glb["Ad_sys"].toarray(), Bd=glb["Bd_sys"].toarray(),
                        update_steps=keep_steps, l_updates=l_up[keep_steps], u_updates=u_up[keep_steps],
                        x1=np.array(glb["plt_x_1"]), x2=np.array(glb["plt_x_2"]), x3=np.array(glb["plt_x_3"]),
                        x4=np.array(glb["plt_x_4"]), u_applied=np.array(glb["plt_u"]),
                        del_u=np.array(glb["plt_del_u"]).ravel(), slack=np.array(glb["plt_s"]),
                        sol_401=sols[401], sol_last=sols[-1], iters=iters,
                        status=np.array([s_["status_val"] for s_ in r["results"]]))
    # ---------------------------------------------------------------- mpc_ (per-stage corridor), mpc__ and the kinematic
    # mpc_increment of the "predictive linearised matrix" scripts.  (Added last: the seeded draws above stay as they were.)
    mkp = _import_ref("Control/MPC/mpc_kinematics_pred_matrix.py", "ref_mpc_kinematics_pred_matrix")
    mip = _import_ref("Control/MPC/mpc_increment_kinematics_pred_matrix.py", "ref_mpc_increment_kinematics_pred_matrix")
    N = 20
    Q = sp.diags([1.0, 1.0, 5.0, 10.0]); QN = sp.diags([10.0, 10.0, 50.0, 50.0]); R = sp.diags([0.1, 0.1])
    umin = np.array([-np.deg2rad(15), -3.]); umax = np.array([np.deg2rad(15), 1.])
    xmin = np.array([-np.inf, -np.inf, -100., -np.pi]); xmax = np.array([np.inf, np.inf, 100., np.pi])
    x = np.array([[0.5], [-0.3], [6.0], [np.deg2rad(20)]]); u = np.array([[0.01], [0.2]])
    A_, B_, C_ = kin.get_kinematics_model(x, u)
    Xr = np.zeros((4, N + 1)); Xr[0] = np.linspace(0.5, 3.5, N + 1); Xr[1] = np.linspace(-0.3, 1.0, N + 1); Xr[2] = 8.0
    lb_x = np.linspace(-1.0, 1.5, N + 1); ub_x = lb_x + 3.0
    lb_y = np.full(N + 1, -2.0); ub_y = np.linspace(1.0, 2.5, N + 1)
    records.clear()
    res = mk.mpc_(A_, B_, C_, x[:, 0], Xr, Q, QN, R, N, lb_x, ub_x, lb_y, ub_y, umin, umax)
    rc = records[0]
    # per-stage linearisations along a rollout with the operating input (what the script's main() feeds mpc__)
    Al, Bl, gl = [], [], []
    xk_ = x.copy()
    for i in range(N + 1):
        a, b, c = kin.get_kinematics_model(xk_, u)
        Al.append(a); Bl.append(b); gl.append(c)
        xk_ = a @ xk_ + b @ u + c
    records.clear()
    res2 = mkp.mpc__(Al, Bl, gl, x[:, 0], Xr, Q, QN, R, N, xmin, xmax, umin, umax)
    rl = records[0]
    del_umin = np.array([-np.deg2rad(2.0), -0.5]); del_umax = np.array([np.deg2rad(2.0), 0.5])
    xmin_t = np.concatenate([xmin, umin]); xmax_t = np.concatenate([xmax, umax])
    records.clear()
    px = np.zeros((6, N + 1)); pdu = np.zeros((2, N + 1))
    with contextlib.redirect_stdout(io.StringIO()):
        px, pdu = mip.mpc_increment(Al, Bl, gl, np.concatenate([x[:, 0], u[:, 0]]), Xr, px, pdu, Q, QN, R, N,
                                    xmin_t, xmax_t, del_umin, del_umax)
    ri = records[0]
    np.savez_compressed(os.path.join(OUT, "qp_kinematic_corridor_predmatrix.npz"), N=N, Ad=A_, Bd=B_, gd=C_[:, 0],
                        x_init=x[:, 0], u_init=u[:, 0], Xr=Xr, Q=Q.diagonal(), QN=QN.diagonal(), R=R.diagonal(),
                        xmin=xmin, xmax=xmax, umin=umin, umax=umax, lb_x=lb_x, ub_x=ub_x, lb_y=lb_y, ub_y=ub_y,
                        corr_l=rc["l"], corr_u=rc["u"], corr_q=rc["q"], corr_x=res.x, corr_iter=res.info.iter,
                        corr_status=res.info.status_val,
                        Ad_list=np.stack(Al), Bd_list=np.stack(Bl), gd_list=np.stack([g[:, 0] for g in gl]),
                        list_A=rl["A"], list_l=rl["l"], list_u=rl["u"], list_q=rl["q"], list_x=res2.x,
                        list_iter=res2.info.iter, list_status=res2.info.status_val,
                        del_umin=del_umin, del_umax=del_umax, xmin_t=xmin_t, xmax_t=xmax_t,
                        inc_A=ri["A"], inc_l=ri["l"], inc_u=ri["u"], inc_q=ri["q"], inc_pred_x=px, inc_pred_du=pdu,
                        inc_iter=ri["results"][0]["iter"], inc_status=ri["results"][0]["status_val"])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("closed-loop iterations per step (H=20, 120 steps): min %d median %d max %d" % (iters20.min(), np.median(iters20), iters20.max()))
    print("closed-loop iterations per step (H=100, 1500 steps): min %d median %d max %d; max |slack| %.3f"
          % (iters.min(), np.median(iters), iters.max(), np.abs(np.array(glb["plt_s"])).max()))


if __name__ == "__main__":
    main()
