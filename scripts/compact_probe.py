"""headline step time over 12 consecutive solves (the learnt compaction point adapts between solves)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
dev = torch.device("cuda", 0)
B = 65536
wl = workloads.lateral_slack_increment(B, seed=7000, dtype=torch.float64)
x0, xr, sp = (torch.as_tensor(v).to(dev) for v in (wl.x0, wl.xr, wl.speed))
ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
ts = []
for i in range(14):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(" ".join("%.2f" % t for t in ts))
os.environ["MPCB_TRACE"] = "1"
ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True); torch.cuda.synchronize()
