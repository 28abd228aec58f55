"""configs[1] (batch 1024 vanilla lateral MPC, one shared linearisation, f64): dense shared-KKT path on / off, the parts of a
step timed separately, and the single-vehicle closed-loop step (update + warm-started solve)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


be = pm.cuda_backend()
dev = torch.device("cuda", 0)
for B in (1024, 1):
    wl = workloads.lateral_vanilla_shared(B, seed=1)
    x0, xr = torch.as_tensor(wl.x0).to(dev), torch.as_tensor(wl.xr).to(dev)
    for dense in (0, 1):
        be.set_option("dense", dense)
        ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
        ms = timed(lambda: ctl.solve_batch(x0, xr, None, want_x=False, reuse=True))
        s = ctl.solver
        n0 = be.launch_count()
        ctl.solve_batch(x0, xr, None, want_x=False, reuse=True)
        nl = be.launch_count() - n0
        ms_solve = timed(lambda: (s.cold_start(), s.solve()))
        ms_upd = timed(lambda: ctl.update_batch(x0 * 0.97))
        it = s.info().iter.double()
        print("B=%d dense=%d: solve_batch %.3f ms (%.3g QP solves/s, %d launches), cold solve alone %.3f ms, update+warm solve+gather %.3f ms, "
              "mean iterations %.1f" % (B, dense, ms, B / (ms * 1e-3), nl, ms_solve, ms_upd, it.mean().item()))
be.set_option("dense", 1)
