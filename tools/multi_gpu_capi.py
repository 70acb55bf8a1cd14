"""Multi-GPU through ONE C-ABI handle (direct_ddp_opts.devices[]): a single host process, one host thread +
stream per device inside the library, host buffers in pinned memory, results D2H straight into the caller's arrays.
    python tools/multi_gpu_capi.py --total 65536 --gpus 1 2 4 8          # strong scaling of BASELINE configs[2]
    python tools/multi_gpu_capi.py --per-gpu 8192 --gpus 1 2 4 8         # weak scaling (8192 trajectories per GPU)
One JSON line per configuration: solves/s end to end (H2D + solve + D2H inside the timed call), max kernel ms over devices."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import HostResult, Solver  # noqa: E402
from direct_b200.problems import ProblemBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=0)
ap.add_argument("--per-gpu", type=int, default=0)
ap.add_argument("--gpus", type=int, nargs="+", default=[1, 2])
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--kind", default="box")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()


def pinned(arr):
    t = torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype, pin_memory=True)
    t.numpy()[...] = arr
    return t


base = None
for n in a.gpus:
    if n > torch.cuda.device_count():
        print(json.dumps({"gpus": n, "skipped": f"only {torch.cuda.device_count()} devices visible"}))
        continue
    B = a.total if a.total else a.per_gpu * n
    pb = make_batch(B, a.knots, a.kind)
    keep = {k: pinned(getattr(pb, k)) for k in ("planes", "nplanes", "durations", "seeds", "x0", "xd")}
    pbp = ProblemBatch(B, a.knots, pb.P_max, *[keep[k].numpy() for k in ("planes", "nplanes", "durations", "seeds", "x0", "xd")],
                       pb.max_vel, pb.max_acc)
    out = HostResult(B, a.knots)
    for name in ("rtn", "infeas_out", "line_failed_out", "iters", "cost", "x_final", "poly_coeff", "bez_coeff", "poly_time", "jerk", "stats"):
        t = pinned(getattr(out, name)); keep["o_" + name] = t; setattr(out, name, t.numpy())
    s = Solver(0, "fp64", devices=list(range(n)))
    s.solve_two_stage(pbp, want_stage0=False, out1=out)   # warm-up: allocations, first touch
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        s.solve_two_stage(pbp, want_stage0=False, out1=out)
        ts.append(time.perf_counter() - t0)
    st = s.stats()
    dt = min(ts)
    line = {"metric": "DDP solves/sec through one C-ABI handle (host buffers, H2D + solve + D2H)", "gpus": n, "batch_total": B,
            "knots": a.knots, "corridor": a.kind, "value": B / dt, "unit": "solves/s", "ms_per_call": dt * 1e3,
            "kernel_ms_max_over_devices": st.kernel_ms, "h2d_bytes": int(st.h2d_bytes), "d2h_bytes": int(st.d2h_bytes),
            "converged_frac": float(np.isin(out.rtn, (1, 2)).mean()), "scaling": "strong" if a.total else "weak"}
    if base is None:
        base = (n, line["value"])
    line["efficiency_vs_first"] = line["value"] / base[1] / ((n / base[0]) if not a.total else (n / base[0]))
    print(json.dumps(line), flush=True)
    s.close()
    del keep, out, pbp, pb
