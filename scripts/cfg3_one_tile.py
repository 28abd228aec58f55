import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads, vehicle_models
N, B = 100, 32
wl = workloads.DynamicWorkload(B, N=N, seed=1)
veh = vehicle_models.Vehicle_Dynamics(dt=wl.dt)
A, Bm, g, _, ld = workloads.rollout_linearisation(veh, wl.x0, wl.u0, N)
Xr = wl.references()
s = pm.BatchSolver(N, 6, 2, wl.Q, wl.QN, wl.R, wl.xmin, wl.xmax, wl.umin, wl.umax, dtype=torch.float64, time_varying=True,
                   stage_reference=True, capacity=B, rho=0.1, eps_abs=1e-4, eps_rel=1e-4, warm_start=False, max_iter=100)
xr = torch.as_tensor(Xr).transpose(1, 2).contiguous()
s.batch = B
s.setup(A, Bm, g, s.to_element_major(wl.x0, B, 6, ld), s.to_element_major(xr, B, (N + 1) * 6, ld), element_major=True)
s.solve(); torch.cuda.synchronize()
print("done", s.info().iter.max().item())
