"""BASELINE.json configs[4] for model (B) (the literal 12-state / 4-input quadrotor DDP): horizon x batch sweep
N in {50,100,200,400} x B in {1k,4k,16k,64k}, fp32 and fp64, device-resident inputs and outputs (no host copies in the timing),
best of `reps` launches -> one markdown table.  NO reference parity for this model.
    python tools/gddp_sweep.py [--reps 2] > profiles/..._gddp_sweep.md"""
import argparse, dataclasses, sys
import torch
sys.path.insert(0, ".")
from direct_b200 import gddp               # noqa: E402
from direct_b200.capi import Solver        # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
fl = gddp.bwd_flops_per_knot(12, 4)
solvers = {p: Solver(0, p) for p in ("fp32", "fp64")}
peaks = {p: s.fma_peak_tflops(p) for p, s in solvers.items()}
print(f"| N | B | fp32 ms | fp32 solves/s | fp32 bwd TFLOP/s | % of fp32 FMA peak ({peaks['fp32']:.1f} TF) | fp64 ms | fp64 solves/s | fp64 bwd TFLOP/s | "
      f"% of fp64 FMA peak ({peaks['fp64']:.1f} TF) | converged fp32 / fp64 |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
stream = torch.cuda.current_stream().cuda_stream
for N in (50, 100, 200, 400):
    for B in (1024, 4096, 16384, 65536):
        gp0 = gddp.make_quad_batch(B, N)
        t = {k: torch.from_numpy(getattr(gp0, k)).to(dev) for k in ("x0", "xg")}
        o = dict(rtn=torch.zeros(B, dtype=torch.int32, device=dev), iters=torch.zeros(B, dtype=torch.int32, device=dev),
                 cost=torch.zeros(B, dtype=torch.float64, device=dev), x=torch.zeros(B, N + 1, 12, dtype=torch.float64, device=dev),
                 u=torch.zeros(B, N, 4, dtype=torch.float64, device=dev), stats=torch.zeros(B, 4, dtype=torch.int64, device=dev))
        oc = gddp.ResultC(*[o[n].data_ptr() for n, _ in gddp.ResultC._fields_])
        row = []
        for prec in ("fp32", "fp64"):
            gp = dataclasses.replace(gp0, tol=1e-5) if prec == "fp32" else gp0
            pc = gddp.problem_struct(gp, t["x0"].data_ptr(), t["xg"].data_ptr(), 0)
            best = None
            for _ in range(a.reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gddp.solve_device(solvers[prec], pc, oc, stream); e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            knots = int(o["stats"][:, 2].sum().item())
            conv = float((o["rtn"] == 1).float().mean().item())
            tf = fl * knots / (best * 1e-3) / 1e12
            row.append((best, B / best * 1e3, tf, tf / peaks[prec] * 100, conv))
        (m32, s32, t32, p32, c32), (m64, s64, t64, p64, c64) = row
        print(f"| {N} | {B} | {m32:.2f} | {s32:.0f} | {t32:.2f} | {p32:.1f} | {m64:.2f} | {s64:.0f} | {t64:.2f} | {p64:.1f} | {c32:.3f} / {c64:.3f} |", flush=True)
        del o, t
        torch.cuda.empty_cache()
for s in solvers.values():
    s.close()
