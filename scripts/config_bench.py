"""Throughput of the other BASELINE configs on one GPU (numbers quoted in DESIGN.md; parity is in tests/):
configs[1] batch 1024 vanilla lateral MPC, one shared linearisation, FP64;
configs[3] batch 8192, H = 100, time-varying combined longitudinal-lateral dynamics model, FP64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads, vehicle_models


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


# configs[1]
wl = workloads.lateral_vanilla_shared(1024, seed=1)
ctl = wl.make_controller(rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
dev = torch.device("cuda", 0)
x0, xr = torch.as_tensor(wl.x0).to(dev), torch.as_tensor(wl.xr).to(dev)
ms = timed(lambda: ctl.solve_batch(x0, xr, None, want_x=False, reuse=True))
it = ctl.solver.info().iter.double()
print("configs[1]: 1024 QPs (vanilla, shared linearisation, f64) in %.2f ms -> %.3g QP solves/s, mean iterations %.1f" % (ms, 1024 / (ms * 1e-3), it.mean().item()))

# configs[3]
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


def run():
    s.setup(A, Bm, g, x_em, xr_em, element_major=True)
    s.cold_start(); s.solve()


ms = timed(run, reps=3)
inf = s.info()
it = inf.iter.double()
print("configs[3]: 8192 QPs (H=100, time-varying dynamics model, f64, rho=0.1) setup+solve in %.1f ms -> %.3g QP solves/s, "
      "mean iterations %.1f, solved %.4f" % (ms, B / (ms * 1e-3), it.mean().item(), (inf.status_val == 1).double().mean().item()))
ms_lin = timed(lambda: workloads.rollout_linearisation(veh, wl.x0, wl.u0, N), reps=3)
print("configs[3]: rollout + linearisation of 8192 x 100 stages (QP build) %.2f ms" % ms_lin)
