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


class PackedResults:
    """The result fields of one rank in ONE contiguous buffer, so that the gather of the solved trajectories is a single
    collective into a buffer preallocated on the root: the solver writes straight into the named views (no packing copy),
    the root reads the gathered fields as per-rank views (no concatenation copy).  Replaces one gather + one torch.cat per
    field (round 1: four of each, +7 ms per step at 8 GPUs for 0.5 GB over NVLink).

    fields: {name: (per-trajectory shape tuple, torch dtype)}; rows = trajectories of this rank; max_rows = largest shard."""

    def __init__(self, fields: dict, rows: int, max_rows: int, device, mode: str = "gather"):
        self.fields, self.rows, self.max_rows, self.device = fields, rows, max_rows, device
        self.offsets, off = {}, 0
        for name, (shape, dtype) in fields.items():
            nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size() * max_rows
            self.offsets[name] = (off, nbytes)
            off += (nbytes + 255) // 256 * 256          # every field 256-byte aligned
        self.nbytes = off
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.root = None
        # mode "allgather" (NCCL only): every rank receives every shard in one all-gather instead of a gather to the root.  Measured
        # with tools/gather_bench.py on two B200s: 0.158 ms (gather) against 0.166 ms (all-gather) for 62 MB per rank - the
        # collective is not what separates the multi-GPU step from the single-GPU one (that is the maximum over the ranks'
        # heavy-tailed kernels), so the default stays the gather.
        self.allgather = dist.is_initialized() and dist.get_backend() == "nccl" and mode == "allgather"
        if dist.is_initialized() and (dist.get_rank() == 0 or self.allgather):
            self.root = torch.empty(dist.get_world_size() * self.nbytes, dtype=torch.uint8, device=device)

    def _view(self, buf, name, rows):
        shape, dtype = self.fields[name]
        off, nbytes = self.offsets[name]
        return buf[off:off + nbytes].view(dtype).view((self.max_rows,) + tuple(shape))[:rows]

    def view(self, name: str) -> torch.Tensor:
        """This rank's tensor of field `name` (rows x shape), a view into the packed buffer."""
        return self._view(self.local, name, self.rows)

    def gather(self, counts: list[int]):
        """One collective.  On the root: {name: [tensor of rank 0, tensor of rank 1, ...]} (views), else None."""
        world = dist.get_world_size()
        bufs = [self.root[r * self.nbytes:(r + 1) * self.nbytes] for r in range(world)] if self.root is not None else None
        if self.allgather:
            dist.all_gather_into_tensor(self.root, self.local)
            if dist.get_rank() != 0:
                return None
        else:
            dist.gather(self.local, bufs, dst=0)
        if self.root is None:
            return None
        return {name: [self._view(bufs[r], name, counts[r]) for r in range(world)] for name in self.fields}


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device) -> np.ndarray:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
