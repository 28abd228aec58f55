"""configs[4]: launch timeline (MPCB_TRACE timestamps) of warm-started closed-loop steps 50 and 150 of a 131072-scenario sweep"""
import sys, os
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
s = ctl.solver
be = s.be
res = ctl.solve_batch(x0, xr, sp, want_x=False)
A, Bm = s._keep["Ad"], s._keep["Bd"]
x_em = s._keep["x_init"]
u = res.u
for k in range(1, 152):
    xn = torch.empty_like(x_em)
    be.check(be.lib.mpcb_plant_step(_dt(torch.float64), B, s.ld, 5, 1, 0, ptr(A), ptr(Bm), ptr(None), ptr(x_em), ptr(u), 20, ptr(xn), be.stream()))
    x_em = xn
    s.update(x_init=x_em, element_major=True)
    if k in (50, 150):
        torch.cuda.synchronize()
        sys.stderr.write("---- step %d\n" % k); sys.stderr.flush()
        os.environ["MPCB_TRACE"] = "1"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve(); e1.record()
    if k in (50, 150):
        torch.cuda.synchronize()
        del os.environ["MPCB_TRACE"]
        it = s.info().iter.cpu().numpy()
        sys.stderr.write("solve %.2f ms, iterations %s\n" % (e0.elapsed_time(e1), dict(zip(*np.unique(it, return_counts=True)))))
    _, _, u = s.solution(want_x=False, want_y=False, want_u=True)
