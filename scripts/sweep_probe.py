"""configs[4] closed-loop sweep on one GPU: per-step timing and the launch schedule (MPCB_TRACE=1)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
wl = workloads.lateral_closed_loop_sweep(B, seed=9000)
x0, xr, sp = (torch.as_tensor(v).to(dev) for v in (wl.x0, wl.xr, wl.speed))
ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
ctl.closed_loop_batch(x0, xr, sp, steps=3, record=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
_, us, its = ctl.closed_loop_batch(x0, xr, sp, steps=steps, record=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
it = its.double()
print("B=%d: %d steps in %.1f ms (%.1f ms per step), mean iterations per step %s, QP solves/s %.3g"
      % (B, steps, dt * 1e3, dt * 1e3 / steps, it.mean(dim=1).cpu().numpy().round(1), B * steps / dt))
print("iteration histogram of the last step:", np.unique(its[-1].cpu().numpy(), return_counts=True))
