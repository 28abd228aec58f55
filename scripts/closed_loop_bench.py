"""BASELINE configs[4] in small: B scenarios x `steps` warm-started MPC steps on one GPU (closed_loop_batch).
Prints scenario-steps (= QP solves) per second.  Not the headline metric — a measurement for DESIGN.md."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
wl = workloads.lateral_slack_increment(a.batch, seed=3, dtype=torch.float64)
ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
ctl.closed_loop_batch(wl.x0, wl.xr, wl.speed, steps=3, record=False)      # warm-up (allocations, re-tile point)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_, us, its = ctl.closed_loop_batch(wl.x0, wl.xr, wl.speed, steps=a.steps, record=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
it = its.double()
print("closed loop: %d scenarios x %d steps in %.1f ms -> %.3g QP solves/s; mean iterations step0 %.1f, later steps %.1f; ms per later step %.2f"
      % (a.batch, a.steps, ms, a.batch * a.steps / (ms * 1e-3), it[0].mean().item(), it[1:].mean().item(), ms / a.steps))
