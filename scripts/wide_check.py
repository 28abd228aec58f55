"""A/B of the 8-lanes-per-QP straggler kernel (option "wide") against the main kernel: same statuses and iteration
counts, solutions to the last bits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
be = pm.cuda_backend()
wl = workloads.lateral_slack_increment(16384, seed=77, dtype=torch.float64)
out = []
for wide in (0, 1):
    be.set_option("wide", wide)
    ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
    n0 = be.launch_count()
    r1 = ctl.solve_batch(wl.x0, wl.xr, wl.speed)
    r1b = ctl.solve_batch(wl.x0, wl.xr, wl.speed)        # second solve: re-tile point learnt
    nl = be.launch_count() - n0
    r2 = ctl.update_batch(wl.x0 * 0.9)
    out.append((r1b.x.clone(), r1b.info.iter.clone(), r1b.info.status_val.clone(), r2.x.clone(), r2.info.iter.clone()))
    print("wide", wide, "launches", nl, "iters", np.unique(r1b.info.iter.cpu().numpy(), return_counts=True))
be.set_option("wide", 1)
a, b = out
print("iter equal:", torch.equal(a[1], b[1]), torch.equal(a[4], b[4]), "status equal:", torch.equal(a[2], b[2]))
print("max |dx| solve:", float((a[0] - b[0]).abs().max()), "rel", float((a[0] - b[0]).abs().max() / a[0].abs().max()))
print("max |dx| warm solve:", float((a[3] - b[3]).abs().max()))

# small batches: the wide schedule runs every iteration between termination tests
for B in (1, 37, 1024):
    wl = workloads.lateral_slack_increment(B, seed=5, dtype=torch.float64)
    out = []
    for wide in (0, 1):
        be.set_option("wide", wide)
        ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
        n0 = be.launch_count()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r1 = ctl.solve_batch(wl.x0, wl.xr, wl.speed); e1.record(); torch.cuda.synchronize()
        r2 = ctl.update_batch(wl.x0 * 0.9)
        out.append((r1.x.clone(), r1.info.iter.clone(), r1.info.status_val.clone(), r2.x.clone(), r2.info.iter.clone()))
        print("B", B, "wide", wide, "launches", be.launch_count() - n0, "solve ms %.2f" % e0.elapsed_time(e1), "max it", int(r1.info.iter.max()))
    be.set_option("wide", 1)
    a, b = out
    print("  iter equal:", torch.equal(a[1], b[1]), torch.equal(a[4], b[4]), "status equal:", torch.equal(a[2], b[2]),
          "max |dx| %.2e / %.2e" % (float((a[0] - b[0]).abs().max()), float((a[3] - b[3]).abs().max())))
