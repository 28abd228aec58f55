"""Why is a bench step occasionally 15-40 ms slower?  Per-step time next to allocator / GC counters."""
import sys, os, gc, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
B = 65536
sets = [workloads.lateral_slack_increment(B, seed=i, dtype=torch.float64) for i in range(4)]
dev = torch.device("cuda", 0)
dev_in = [(torch.as_tensor(w.x0).to(dev), torch.as_tensor(w.xr).to(dev), torch.as_tensor(w.speed).to(dev)) for w in sets]
ctl = sets[0].make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
for i in range(24):
    x0, xr, sp = dev_in[i % 4]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st0 = torch.cuda.memory_stats().get("num_device_alloc", 0); g0 = gc.get_count()
    r = ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
    it = r.info.iter.clone()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    st1 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    print("step %2d %.2f ms  device_allocs +%d  gc %s -> %s  max_it %d" % (i, dt, st1 - st0, g0, gc.get_count(), int(it.max())))
