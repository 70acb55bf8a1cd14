"""Cost of bringing the solved trajectories of every rank together (the collective inside bench.py's multi-GPU step):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gather_bench.py [--batch 4096] [--knots 100]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from direct_b200 import dist as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
N = a.knots
fields = {"poly_time": ((N,), torch.float64), "bez_coeff": ((N, 18), torch.float64), "rtn": ((), torch.int32), "cost": ((), torch.float64)}
for mode in ("gather", "allgather"):
    p = D.PackedResults(fields, a.batch, a.batch, dev, mode=mode)
    p.view("cost").fill_(float(rank + 1))
    for _ in range(3):
        p.gather([a.batch] * world)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = p.gather([a.batch] * world)
    e1.record(); torch.cuda.synchronize()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / 10, dev)
    if rank == 0:
        ok = all(float(out["cost"][r][0]) == r + 1 and float(out["cost"][r][-1]) == r + 1 for r in range(world))
        print(f"{world} ranks, {p.nbytes / 1e6:.1f} MB per rank, mode {mode!r} ({'all-gather' if p.allgather else 'gather to root'}): "
              f"{ms:.3f} ms per collective, contents {'ok' if ok else 'WRONG'}", flush=True)
dist.destroy_process_group()
