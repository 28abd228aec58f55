"""per-GPU batch sizes of the strong-scaling record (65536 / N): default schedule vs the CTA-per-tile kernel ("cta" = 2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
be = pm.cuda_backend()
dev = torch.device("cuda", 0)
for B in [int(v) for v in os.environ.get('PROBE_B', '2048,8192,16384,32768,65536').split(',')]:
    wl = workloads.lateral_slack_increment(B, seed=7000, dtype=torch.float64)
    x0, xr, sp = (torch.as_tensor(v).to(dev) for v in (wl.x0, wl.xr, wl.speed))
    for cta in (1, 2):
        be.set_option("cta", cta)
        ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
        for _ in range(3):
            ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("B=%d cta=%d: %.2f ms per step -> %.3g QP solves/s" % (B, cta, np.median(ts), B / (np.median(ts) * 1e-3)))
be.set_option("cta", 1)
