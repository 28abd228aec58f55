import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
be = pm.cuda_backend()
wl = workloads.lateral_slack_increment(65536, seed=3, dtype=torch.float64)
ctls = {}
for wide in (0, 1):
    be.set_option("wide", wide)
    ctls[wide] = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    ctls[wide].closed_loop_batch(wl.x0, wl.xr, wl.speed, steps=3, record=False)
for rep in range(4):
    for wide in (0, 1):
        be.set_option("wide", wide)
        torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = be.launch_count()
        _, us, its = ctls[wide].closed_loop_batch(wl.x0, wl.xr, wl.speed, steps=10, record=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print("rep %d wide %d: %.1f ms per step, launches %d, mean it %.1f max it %d" % (rep, wide, dt / 10, be.launch_count() - n0, its.double().mean().item(), int(its.max())))
