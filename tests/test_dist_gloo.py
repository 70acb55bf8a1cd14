"""world_size-2 gloo test of the multi-GPU host logic (sharding, option broadcast, ragged gather).
The per-rank 'solve' is the CPU emulation of the kernel, so the gathered result can be compared with a
single-process solve of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from direct_b200.dist import shard_range


def test_shard_range_covers_everything():
    for total, world in ((4096, 8), (10, 4), (3, 8), (65536, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, N, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from direct_b200 import dist as D
    from direct_b200.problems import STAGE0, STAGE1, make_batch
    from tools import emu_py
    opts = None
    if rank == 0:
        opts = dict(w_snap0=1.0, w_terminal0=1.0, w_time0=1.0, iter_max0=50, w_snap=1.0, w_terminal=100.0, w_time=20.0,
                    iter_max=100, time_power=2, max_vel=2.0, max_acc=2.0)
    opts = D.broadcast_options(opts, "cpu")
    assert opts["iter_max"] == 100 and opts["w_time"] == 20.0
    lo, hi = D.shard_range(total, rank, world)
    pb = make_batch(hi - lo, N, "box", first=lo)
    s0 = dict(w_snap=opts["w_snap0"], w_terminal=opts["w_terminal0"], w_time=opts["w_time0"], iter_max=opts["iter_max0"])
    s1 = dict(w_snap=opts["w_snap"], w_terminal=opts["w_terminal"], w_time=opts["w_time"], iter_max=opts["iter_max"])
    _, r = emu_py.two_stage_batch(pb, stage0=s0, stage1=s1, time_power=opts["time_power"])
    local = dict(rtn=torch.from_numpy(r.rtn), cost=torch.from_numpy(r.cost), bez=torch.from_numpy(r.bez_coeff),
                 time=torch.from_numpy(r.poly_time))
    counts = [b - a for a, b in (D.shard_range(total, k, world) for k in range(world))]
    g = D.gather_results(local, counts)
    # the same through the packed single-collective path (what bench.py uses): fields are views of one buffer per rank
    pk = D.PackedResults({"rtn": ((), torch.int32), "cost": ((), torch.float64), "bez": ((N, 18), torch.float64),
                          "time": ((N,), torch.float64)}, hi - lo, max(counts), "cpu")
    for name, tsr in local.items():
        pk.view(name).copy_(tsr)
    gp = pk.gather(counts)
    if rank == 0:
        for name in local:
            assert torch.equal(torch.cat(gp[name], dim=0), g[name]), name
    else:
        assert gp is None
    t = D.max_over_ranks(float(rank + 1), "cpu")
    assert t == float(world)
    if rank == 0:
        q.put({k: v.numpy() for k, v in g.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_solve_gather_matches_single_process():
    from direct_b200.problems import make_batch
    from tools import emu_py
    emu_py.lib()  # build before forking
    total, N, world = 5, 6, 2   # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    _, ref = emu_py.two_stage_batch(make_batch(total, N, "box", first=0))
    assert np.array_equal(got["rtn"], ref.rtn) and np.array_equal(got["cost"], ref.cost)
    assert np.array_equal(got["bez"], ref.bez_coeff) and np.array_equal(got["time"], ref.poly_time)
