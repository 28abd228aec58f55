"""configs[4]: where does a closed-loop step's time go over a long sweep?  (per-50-step wall time, iteration histogram)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
from python_mpc_b200._lib import ptr
from python_mpc_b200.vehicle_models import _dt
B = 131072
dev = torch.device("cuda", 0)
wl = workloads.lateral_closed_loop_sweep(B, seed=9000)
x0, xr, sp = (torch.as_tensor(v).to(dev) for v in (wl.x0, wl.xr, wl.speed))
ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
ctl.closed_loop_batch(x0, xr, sp, steps=3, record=False)
s = ctl.solver
be = s.be
res = ctl.solve_batch(x0, xr, sp, want_x=False)
A, Bm = s._keep["Ad"], s._keep["Bd"]
x_em = s._keep["x_init"]
u = res.u
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else 'plain'
us, its = [], []
t_last = time.perf_counter()
for k in range(1, 201):
    xn = torch.empty_like(x_em)
    be.check(be.lib.mpcb_plant_step(_dt(torch.float64), B, s.ld, 5, 1, 0, ptr(A), ptr(Bm), ptr(None), ptr(x_em), ptr(u), 20, ptr(xn), be.stream()))
    x_em = xn
    s.update(x_init=x_em, element_major=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve(); e1.record()
    _, _, u = s.solution(want_x=False, want_y=False, want_u=True)
    info = s.info()
    if mode in ('keep', 'keep_u'):
        us.append(u[:, 0, :].clone())
    if mode in ('keep', 'keep_it'):
        its.append(info.iter)
    if k % 25 == 0:
        torch.cuda.synchronize()
        now = time.perf_counter()
        it = info.iter.cpu().numpy()
        print("steps ..%d: %.1f ms per step (last solve() alone %.1f ms), iterations %s" % (k, (now - t_last) * 1e3 / 25, e0.elapsed_time(e1), dict(zip(*np.unique(it, return_counts=True)))))
        t_last = time.perf_counter()
