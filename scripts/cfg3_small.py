"""latency floor of the CTA-per-tile kernel: a small time-varying batch (few tiles, GPU mostly idle)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads, vehicle_models
N = 100
for B in (32, 1024, 4096):
    wl = workloads.DynamicWorkload(B, N=N, seed=1)
    veh = vehicle_models.Vehicle_Dynamics(dt=wl.dt)
    A, Bm, g, _, ld = workloads.rollout_linearisation(veh, wl.x0, wl.u0, N)
    Xr = wl.references()
    s = pm.BatchSolver(N, 6, 2, wl.Q, wl.QN, wl.R, wl.xmin, wl.xmax, wl.umin, wl.umax, dtype=torch.float64, time_varying=True,
                       stage_reference=True, capacity=B, rho=0.1, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
    xr = torch.as_tensor(Xr).transpose(1, 2).contiguous()
    s.batch = B
    x_em = s.to_element_major(wl.x0, B, 6, ld); xr_em = s.to_element_major(xr, B, (N + 1) * 6, ld)
    for _ in range(2):
        s.setup(A, Bm, g, x_em, xr_em, element_major=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(); e1.record(); torch.cuda.synchronize()
    it = s.info().iter.cpu().numpy()
    ms = e0.elapsed_time(e1)
    print("B=%d: solve %.1f ms, max iterations %d -> %.1f us per iteration of the slowest tile, %.2f us per stage sweep"
          % (B, ms, it.max(), 1e3 * ms / it.max(), 1e3 * ms / it.max() / (2 * (N + 1))))
