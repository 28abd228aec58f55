"""Small solves through every ADMM code path, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
be = pm.cuda_backend()
be.set_option("retile_min_batch", 64)
for B, slack, inc in ((300, True, True), (70, False, False), (33, True, False)):
    wl = workloads.LateralWorkload(B, 20, slack, inc, 5, torch.float64)
    if not slack:
        wl.x0[1, 3] = 14.0                      # one primal-infeasible QP: certificate sweep + NaN gather
    ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    r = ctl.solve_batch(wl.x0, wl.xr, wl.speed)
    r2 = ctl.update_batch(wl.x0 * 0.9)
    torch.cuda.synchronize()
    print(B, slack, inc, np.unique(r.info.status_val.cpu().numpy(), return_counts=True), float(r2.info.iter.double().mean()))
print("done")
