"""How long do NVML queries take while the GPU is busy, and do they delay a solve?"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import python_mpc_b200 as pm
from python_mpc_b200 import workloads
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
B = 65536
wl = workloads.lateral_slack_increment(B, seed=1, dtype=torch.float64)
dev = torch.device("cuda", 0)
x0, xr, sp = (torch.as_tensor(a).to(dev) for a in (wl.x0, wl.xr, wl.speed))
ctl = wl.make_controller(capacity=B, rho=5.0, eps_abs=1e-4, eps_rel=1e-4, warm_start=False)
for _ in range(3): ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
torch.cuda.synchronize()
stop = False; log = []
def poll(which, period):
    while not stop:
        t0 = time.perf_counter()
        if which == "clock": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        else: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        log.append((time.perf_counter() - t0) * 1e3)
        time.sleep(period)
for which, period in (("none", 0), ("clock", 0.02), ("reasons", 0.02), ("clock", 0.002)):
    stop = False; log = []
    th = None
    if which != "none":
        th = threading.Thread(target=poll, args=(which, period), daemon=True); th.start()
    ts = []
    for i in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    stop = True
    if th: th.join()
    print("%-8s period %.3f: steps %s | query ms: n=%d max %.2f mean %.2f" % (which, period, " ".join("%.1f" % t for t in ts), len(log), max(log) if log else 0, sum(log) / len(log) if log else 0))
