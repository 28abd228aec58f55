#!/usr/bin/env python
"""Headline benchmark: lateral-MPC QP solves/sec (BASELINE.json metric).

Workload (config.workload): configs[2] of BASELINE.json — a batch of 65536 soft-constraint +
incremental (slack + delta-u) lateral MPC QPs, H = 20, every QP linearised at its own vehicle speed,
OSQP-equivalent ADMM with fixed rho/sigma/alpha, adaptive_rho and polish off, eps_abs = eps_rel = 1e-4.
One "step" = one pass of the hot path over one synthetic batch:
    QP build (discretise A,B per speed, delta-u augmentation, layout) -> Ruiz scaling + cached KKT
    factorisation -> ADMM to convergence -> gather of the control sequences.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

value  : whole-job QP solves/sec with the inputs already resident in HBM
e2e    : the same through the public API (LateralMPC.solve_batch) from pinned HOST buffers, H2D of the
         inputs and D2H of the control sequences inside the timed region
--impl reference : the reference's per-QP OSQP loop restated in C (oracle/osqp_admm.c), all host
         threads, on a bounded sample of the same workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lateral-MPC QP solves/sec (batch 65536, H=20)"
UNIT = "QP solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="QPs per GPU (weak scaling)")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--rho", type=float, default=5.0)
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="QPs in the cpu_baseline / reference sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": "BASELINE configs[2]: batch %d slack+delta-u lateral MPC QPs per GPU, H=%d, per-QP speed linearisation"
                        % (a.batch, a.horizon),
            "batch_per_gpu": a.batch, "global_batch": a.batch * n_gpus, "horizon": a.horizon,
            "nvar": (a.horizon + 1) * 5 * 2 + a.horizon, "ncon": 2 * (a.horizon + 1) * 5 + a.horizon,
            "rho": a.rho, "sigma": 1e-6, "alpha": 1.6, "eps_abs": a.eps, "eps_rel": a.eps, "max_iter": 4000,
            "check_termination": 25, "scaling": 10, "adaptive_rho": False, "polish": False, "warm_start": False,
            "l2": "per-step working set (~1.6 GB) exceeds the 126 MB L2; a different random batch every step",
            "parallelism": "independent QPs sharded across %d GPU(s); NCCL all_gather of control sequences only" % n_gpus}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port, timed on the host cores (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_solves_per_sec(a, sample, seed, repeats=1):
    import torch
    from oracle import c_oracle, workload_qp
    from python_mpc_b200 import workloads
    wl = workloads.lateral_slack_increment(sample, N=a.horizon, seed=seed, dtype=torch.float64)
    Pu, A0, Pv, q, Av, l, u, perm = workload_qp.lateral_batch_csc(wl)
    best = None
    ncores = len(os.sched_getaffinity(0))        # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    for _ in range(repeats):
        t0 = time.perf_counter()
        x, y, it, st, used = c_oracle.solve_batch(Pu, A0, Pv, q, Av, l, u, perm=perm, nthreads=ncores, rho=a.rho,
                                                  eps_abs=a.eps, eps_rel=a.eps, max_iter=4000)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, used, float(it.mean()), float((st == 1).mean()), best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = a.cpu_sample or 16384
    times = []
    info = None
    for i in range(a.warmup + a.steps):
        v, cores, iters, solved, dt = cpu_solves_per_sec(a, sample, seed=1000 + i)
        if i >= a.warmup:
            times.append(dt)
        info = (cores, iters, solved)
    tot = sum(times)
    value = sample * a.steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info[0], "kind": "port",
                             "sample": "%d QPs of the same workload per step (oracle/osqp_admm.c, OpenMP over QPs); "
                                       "mean %.1f ADMM iterations, %.3f solved" % (sample, info[1], info[2])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled in-process every ~20 ms (the
    timed region is a few steps of ~25 ms, too short for `nvidia-smi -lms`); nvidia-smi is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.t, self.h, self.nv = index, [], False, None, None, None
        self.sm_max, self.source = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ent = vis.split(",")[index].strip()
                phys = int(ent) if ent.isdigit() else None
                if phys is None:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(ent.encode() if hasattr(ent, "encode") else ent)
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv, self.source = pynvml, "nvml"
        except Exception:
            self.h = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(0.02)       # NVML queries take the driver lock: polled sparingly so that they do not perturb the steps

    def _poll_smi(self):
        names = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().splitlines()[0].split(",")]
                mask = 0
                for bit, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.sm_max = float(r[1])
                self.rows.append((time.perf_counter(), float(r[0]), mask))
            except Exception:
                time.sleep(0.05)

    def start(self):
        if self.h is None:
            self.source = "nvidia-smi"
        self.t = threading.Thread(target=self._poll_nvml if self.h is not None else self._poll_smi, daemon=True)
        self.t.start()

    def window(self, t0, t1):
        """Summary of the samples taken inside [t0, t1] (perf_counter times bracketing the timed region)."""
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = sorted(n for b, n in self.REASONS.items() if mask & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "source": self.source}

    def stop(self):
        self.stop_flag = True
        if self.t is not None:
            self.t.join(timeout=6)


def run_ours(a):
    import torch
    import torch.distributed as dist
    import python_mpc_b200 as pm
    from python_mpc_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    esz = 4 if a.dtype == "f32" else 8
    B, N = a.batch, a.horizon
    be = pm.cuda_backend()
    nsets = a.warmup + a.steps
    # a fresh random batch per step (per rank); resident in HBM for `value`, pinned on the host for `e2e`
    sets = [workloads.lateral_slack_increment(B, N=N, seed=100 * rank + i, dtype=dtype) for i in range(min(nsets, 4))]
    dev_in = [(torch.as_tensor(w.x0).to(dev, dtype), torch.as_tensor(w.xr).to(dev, dtype),
               torch.as_tensor(w.speed).to(dev, dtype)) for w in sets]
    host_in = [(torch.as_tensor(w.x0).to(dtype).pin_memory(), torch.as_tensor(w.xr).to(dtype).pin_memory(),
                torch.as_tensor(w.speed).to(dtype).pin_memory()) for w in sets]
    ctl = sets[0].make_controller(capacity=B, rho=a.rho, eps_abs=a.eps, eps_rel=a.eps, warm_start=False)
    gathered = [torch.empty((B, N, 1), device=dev, dtype=dtype) for _ in range(world)] if world > 1 else None

    solve_ev = []

    def step(i, timed_events=False):
        x0, xr, sp = dev_in[i % len(dev_in)]
        res = ctl.solve_batch(x0, xr, sp, want_x=False, reuse=True)
        if world > 1:
            dist.all_gather(gathered, res.u.contiguous())
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up (NVML's first queries take milliseconds); only the samples that fall
    # inside the timed region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(a.warmup, 1)):        # (at least one untimed step: allocations, the learnt re-tile point)
        res = step(i)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    # per-step solver info goes into buffers allocated BEFORE the timed region: a fresh allocation inside it can make the
    # caching allocator call cudaMalloc, which synchronises the device (seen as one 5-15 ms slower step)
    it_buf = torch.empty((a.steps, B), device=dev, dtype=res.info.iter.dtype)
    st_buf = torch.empty((a.steps, B), device=dev, dtype=res.info.status_val.dtype)
    step(0)                                  # one more untimed step after the allocations above
    barrier()
    launches0 = be.launch_count()
    wall0 = time.perf_counter()
    ev[0].record()
    for i in range(a.steps):
        res = step(a.warmup + i)
        it_buf[i].copy_(res.info.iter); st_buf[i].copy_(res.info.status_val)
        ev[i + 1].record()
    barrier()
    wall1 = time.perf_counter()
    launches = be.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    total_ms = ev[0].elapsed_time(ev[a.steps])
    clocks = None
    if rank == 0:
        clocks = sampler.window(wall0, wall1)
    iters = it_buf.reshape(-1).double()
    solved = (st_buf == 1).double().mean().item()
    mean_iter = iters.mean().item()

    # ---- dominant kernel alone (the ADMM loop): CUDA events on the launching stream
    s = ctl.solver
    admm_ms = []
    for i in range(max(2, min(a.steps, 5))):
        x0, xr, sp = dev_in[i % len(dev_in)]
        ctl.solve_batch(x0, xr, sp, want_x=False)          # sets the batch up
        s.cold_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); s.solve(); e1.record()
        torch.cuda.synchronize()
        admm_ms.append(e0.elapsed_time(e1))
    admm_avg_ms = float(np.mean(admm_ms))
    warp_iters = s.info().iter.double().reshape(-1, 32).max(dim=1).values.sum().item() * 32   # iterations the warps executed

    # ---- e2e: public API from pinned host buffers, copies inside the timed region
    out_host = torch.empty((B, N, 1), dtype=dtype).pin_memory()

    def e2e_step(i):
        hx0, hxr, hsp = host_in[i % len(host_in)]
        r = ctl.solve_batch(hx0.to(dev, non_blocking=True), hxr.to(dev, non_blocking=True),
                            hsp.to(dev, non_blocking=True), want_x=False, reuse=True)
        out_host.copy_(r.u, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(a.steps):
        e2e_step(i)
    t1.record(); barrier()
    e2e_ms = t0.elapsed_time(t1)
    if rank == 0:
        sampler.stop()

    tms = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = tms.tolist()

    if rank == 0:
        from python_mpc_b200 import roofline
        bytes_qp_iter = roofline.admm_bytes_per_qp_iteration(N, 5, 1, True, esz)
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak = peaks.get("hbm_gbs")
        if peak is None:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = bytes_qp_iter * B * mean_iter / (admm_avg_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the two admm_tma_kernel launches of one solve, from the
        # `ncu --set full` capture profiles/r1n_admm_ncu_raw.csv (same command, default workload)
        default_cfg = (B == 65536 and N == 20 and a.dtype == "f64" and a.rho == 5.0 and a.eps == 1e-4)
        traffic = (104.45e9 + 24.72e9) + 0.04e9 if default_cfg else None
        traffic_src = "profiles/r1n_admm_ncu_raw.csv" if default_cfg else None
        cfg = workload_config(a, world)
        ws_mb = s.be.lib.mpcb_workspace_bytes(s._h) / 1e6
        cfg["l2"] = "per-step working set %.0f MB exceeds the 126 MB L2; a different random batch every step" % ws_mb
        cfg.update({"mean_admm_iterations": mean_iter, "fraction_solved": solved})
        line = {"metric": METRIC, "value": B * world * a.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": cfg,
                "p50_batch_latency_ms": float(np.median(step_ms)), "step_ms": [round(v, 2) for v in step_ms],
                "clocks": clocks,
                "e2e": {"value": B * world * a.steps / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": B * (5 + 4 + 1) * esz, "d2h_bytes_per_step": B * N * esz},
                "gpu_launches": launches,
                "roofline": {"kernel": "admm_tma_kernel (ADMM loop: phase-1 launch; + admm_wide_kernel and the tested iteration of the stragglers)",
                             "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_unit": "bytes per solve (the admm_tma_kernel launches: phase 1 + the stragglers' tested iteration)",
                             "traffic_source": traffic_src, "peak_source": peak_src,
                             "algorithmic_bytes_per_qp_iteration": bytes_qp_iter,
                             "algorithmic_bytes_per_solve": bytes_qp_iter * B * mean_iter,
                             "avg_launch_ms": admm_avg_ms,
                             "note": "algorithmic bytes = 164 record elements per stage and iteration x the iterations each "
                                     "QP needs (termination sweeps, certificate sweeps and the old-state copies are not "
                                     "counted); a warp streams its tile's records until its slowest lane converges "
                                     "(%.2fx the needed lane-iterations without re-tiling), which is why unconverged QPs "
                                     "are re-tiled.  ncu of the phase-1 launch: 5.99 TB/s of DRAM traffic = 0.93 of the "
                                     "measured copy bandwidth" % (warp_iters / (B * mean_iter))}}
        if not a.no_cpu_baseline:
            sample = a.cpu_sample or 16384
            v, cores, cit, csolved, dt = cpu_solves_per_sec(a, sample, seed=4242, repeats=2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d QPs of the same workload (oracle/osqp_admm.c, OpenMP over QPs, %.1f s); "
                                              "mean %.1f ADMM iterations, %.3f solved" % (sample, dt, cit, csolved)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
