"""Small solves through every ADMM code path, meant to be run under compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py [main|dense|cta|all]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads, vehicle_models
what = sys.argv[1] if len(sys.argv) > 1 else "all"
be = pm.cuda_backend()
if what in ("main", "all"):
    be.set_option("retile_min_batch", 64)
    for B, slack, inc in ((300, True, True), (70, False, False), (33, True, False)):
        wl = workloads.LateralWorkload(B, 20, slack, inc, 5, torch.float64)
        if not slack:
            wl.x0[1, 3] = 14.0                      # one primal-infeasible QP: certificate sweep + NaN gather
        ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
        r = ctl.solve_batch(wl.x0, wl.xr, wl.speed)
        xmin = wl.xmin.copy(); xmin[3] = 1.0
        ctl.update_bounds(xmin=xmin)                # bound update: per-QP refactor decision on the device
        r2 = ctl.update_batch(wl.x0 * 0.9)
        torch.cuda.synchronize()
        print("main", B, slack, inc, np.unique(r.info.status_val.cpu().numpy(), return_counts=True), float(r2.info.iter.double().mean()))
    be.set_option("retile_min_batch", 4096)
if what in ("dense", "all"):
    for slack, inc, B in ((False, False, 19), (True, True, 9)):
        wl = workloads.LateralWorkload(B, 20, slack, inc, 6, torch.float64, shared_speed=8.3128334)
        if not slack:
            wl.x0[1, 3] = 14.0
        ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
        r = ctl.solve_batch(wl.x0, wl.xr, wl.speed)
        r2 = ctl.update_batch(wl.x0 * 0.9)
        torch.cuda.synchronize()
        print("dense", B, slack, inc, np.unique(r.info.status_val.cpu().numpy(), return_counts=True), float(r2.info.iter.double().mean()))
if what in ("cta", "all"):
    B, N = 70, 12
    wl = workloads.DynamicWorkload(B, N=N, seed=4)
    veh = vehicle_models.Vehicle_Dynamics(dt=wl.dt)
    A, Bm, g, _, ld = workloads.rollout_linearisation(veh, wl.x0, wl.u0, N)
    xr = torch.as_tensor(wl.references()).transpose(1, 2).contiguous()
    s = pm.BatchSolver(N, 6, 2, wl.Q, wl.QN, wl.R, wl.xmin, wl.xmax, wl.umin, wl.umax, dtype=torch.float64, time_varying=True,
                       stage_reference=True, capacity=B, rho=0.1, eps_abs=1e-3, eps_rel=1e-3, warm_start=True, max_iter=200)
    s.batch = B
    s.setup(A, Bm, g, s.to_element_major(wl.x0, B, 6, ld), s.to_element_major(xr, B, (N + 1) * 6, ld), element_major=True)
    s.solve(); s.solve()
    torch.cuda.synchronize()
    print("cta tv", np.unique(s.info().status_val.cpu().numpy(), return_counts=True))
    be.set_option("cta", 2)
    wl = workloads.LateralWorkload(50, 20, True, True, 5, torch.float64)
    ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    r = ctl.solve_batch(wl.x0, wl.xr, wl.speed); r2 = ctl.update_batch(wl.x0 * 0.9)
    torch.cuda.synchronize()
    be.set_option("cta", 1)
    print("cta lateral", np.unique(r.info.status_val.cpu().numpy(), return_counts=True))
print("done")
