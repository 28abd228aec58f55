"""Sharding of the independent QP batch across the GPUs of one box (one process per GPU).

The path has no data-path collective: every rank solves its own contiguous slice of the batch; the only
exchange is the final all_gather of the control sequences (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  BASELINE.json north_star (3)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced slice [lo, hi) of `total` QPs owned by `rank` (first total % world ranks get one more)."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_controls(u_local, total, group=None):
    """all_gather of the per-rank control sequences (B_local, N, nu) -> (total, N, nu) on every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return u_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    if len(set(counts)) == 1:
        out = torch.empty((total,) + tuple(u_local.shape[1:]), device=u_local.device, dtype=u_local.dtype)
        dist.all_gather_into_tensor(out, u_local.contiguous(), group=group)
        return out
    # ragged split: pad every slice to the largest one (collectives need equal sizes), trim after the exchange
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(u_local.shape[1:]), device=u_local.device, dtype=u_local.dtype)
    pad[:counts[rank]] = u_local
    out = torch.empty((world * cmax,) + tuple(u_local.shape[1:]), device=u_local.device, dtype=u_local.dtype)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * cmax:r * cmax + counts[r]] for r in range(world)], dim=0)


def solve_sharded(controller, states, references, speeds, group=None):
    """Every rank passes the GLOBAL batch description; solves its slice; returns (u_global, local BatchResult)."""
    total = states.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(total, rank, world)
    res = controller.solve_batch(states[lo:hi], references[lo:hi], None if speeds is None else speeds[lo:hi], want_x=False)
    return gather_controls(res.u, total, group), res
