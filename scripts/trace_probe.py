"""launch-by-launch timeline (MPCB_TRACE timestamps) of one headline solve and of the per-GPU shares of the strong record"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
be = pm.cuda_backend()
dev = torch.device("cuda", 0)
for B in (65536, 32768, 16384):
    wl = workloads.lateral_slack_increment(B, seed=7000, dtype=torch.float64)
    x0, xr, sp = (torch.as_tensor(v).to(dev) for v in (wl.x0, wl.xr, wl.speed))
    ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
    for _ in range(3):
        ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
    torch.cuda.synchronize()
    sys.stderr.write("---- B=%d\n" % B); sys.stderr.flush()
    os.environ["MPCB_TRACE"] = "1"
    ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
    torch.cuda.synchronize()
    del os.environ["MPCB_TRACE"]
