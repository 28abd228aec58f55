"""N > 1 path on CPU: world_size-2 gloo processes, each solving its slice of the batch through the (emulated)
kernels, all_gather of the control sequences — checked against a single-process solve of the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from python_mpc_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, emu_path, out):
    import ctypes
    sys.path.insert(0, ROOT)
    from python_mpc_b200 import _lib, sharding as sh, vehicle_models, workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = _lib.Backend(ctypes.CDLL(emu_path), "cpu")
    wl = workloads.lateral_slack_increment(total, seed=17, dtype=torch.float64)
    ctl = wl.make_controller(vehicle=vehicle_models.Vehicle_Lateral(_backend=be), _backend=be, rho=5.0, eps_abs=1e-4,
                             eps_rel=1e-4, warm_start=False)
    u, local = sh.solve_sharded(ctl, wl.x0, wl.xr, wl.speed)
    lo, hi = sh.shard_range(total, rank, world)
    assert local.u.shape[0] == hi - lo
    if rank == 0:
        full = ctl.solve_batch(wl.x0, wl.xr, wl.speed, want_x=False)
        out.put(float((full.u - u).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_two_rank_gloo_shard_and_gather(emu_backend, total):
    import __graft_entry__ as g
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 400) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, g.build_emu(), out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() == 0.0
