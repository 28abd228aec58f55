"""one solve of configs[3] (8192 QPs, H = 100, time-varying dynamics model) for ncu: profiles/r2*_cta_*"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads, vehicle_models
B, N = 8192, 100
wl = workloads.DynamicWorkload(B, N=N, seed=1)
veh = vehicle_models.Vehicle_Dynamics(dt=wl.dt)
A, Bm, g, _, ld = workloads.rollout_linearisation(veh, wl.x0, wl.u0, N)
Xr = wl.references()
s = pm.BatchSolver(N, 6, 2, wl.Q, wl.QN, wl.R, wl.xmin, wl.xmax, wl.umin, wl.umax, dtype=torch.float64, time_varying=True,
                   stage_reference=True, capacity=B, rho=0.1, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
xr = torch.as_tensor(Xr).transpose(1, 2).contiguous()
s.batch = B
x_em = s.to_element_major(wl.x0, B, 6, ld); xr_em = s.to_element_major(xr, B, (N + 1) * 6, ld)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    s.setup(A, Bm, g, x_em, xr_em, element_major=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve(); e1.record(); torch.cuda.synchronize()
it = s.info().iter.cpu().numpy()
tmax = it.reshape(-1, 32).max(axis=1)
print("solve %.1f ms; iterations mean %.1f max %d; per-tile max: mean %.1f; histogram %s" % (e0.elapsed_time(e1), it.mean(), it.max(), tmax.mean(), np.unique(it, return_counts=True)))
