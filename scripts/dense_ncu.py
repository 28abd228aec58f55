"""one cold solve of configs[1] (1024 vanilla lateral QPs, one linearisation) for ncu: profiles/r2*_dense_*"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
wl = workloads.lateral_vanilla_shared(1024, seed=1)
ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
for _ in range(2):
    r = ctl.solve_batch(wl.x0, wl.xr, None, want_x=False)
torch.cuda.synchronize()
print("iterations: mean %.1f max %d" % (r.info.iter.double().mean().item(), int(r.info.iter.max())))
