"""Data-parallel sharding of a trajectory batch over ranks (one process per GPU, torch.distributed).

The path shards trivially: trajectories are independent, so there is NO collective
on the data path.  The only exchanges are
  * a broadcast (root 0) of the solver options / stage weights, so every rank solves with identical
    parameters, and
  * a gather (to root 0) of the solved trajectories (segment times + Bezier control points + status).
Works with backend "nccl" (GPU tensors) and "gloo" (CPU tensors; used by the world_size-2 CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

OPT_KEYS = ("w_snap0", "w_terminal0", "w_time0", "iter_max0", "w_snap", "w_terminal", "w_time", "iter_max",
            "time_power", "max_vel", "max_acc")


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_options(opts: dict | None, device) -> dict:
    """Rank 0's options win on every rank (ncclBroadcast of a small fp64 vector)."""
    vec = torch.zeros(len(OPT_KEYS), dtype=torch.float64, device=device)
    if dist.get_rank() == 0:
        vec.copy_(torch.tensor([float(opts[k]) for k in OPT_KEYS], dtype=torch.float64))
    dist.broadcast(vec, src=0)
    out = {k: float(v) for k, v in zip(OPT_KEYS, vec.cpu().tolist())}
    for k in ("iter_max0", "iter_max", "time_power"):
        out[k] = int(out[k])
    return out


def gather_results(local: dict[str, torch.Tensor], counts: list[int]) -> dict[str, torch.Tensor] | None:
    """Gather per-trajectory result tensors (leading dim = local batch) to rank 0, in rank order.
    Shards may be ragged (counts[r] rows on rank r): tensors are padded to the largest shard for the
    collective and trimmed on the root."""
    rank, world = dist.get_rank(), dist.get_world_size()
    m = max(counts)
    out = {} if rank == 0 else None
    for name, t in local.items():
        pad = t
        if t.shape[0] < m:
            pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad.contiguous(), bufs, dst=0)
        if rank == 0:
            out[name] = torch.cat([b[: counts[r]] for r, b in enumerate(bufs)], dim=0)
    return out


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device) -> np.ndarray:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
