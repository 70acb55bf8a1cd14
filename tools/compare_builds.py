"""Do two builds of the library return the same bits?  Each library is loaded in its own process.
    python tools/compare_builds.py LIB_A LIB_B [--batch 512] [--knots 60]"""
import argparse
import os
import subprocess
import sys

import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="*")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--knots", type=int, default=60)
ap.add_argument("--kind", default="box")
ap.add_argument("--dump", default="")
a = ap.parse_args()
FIELDS = ("rtn", "iters", "infeas_out", "cost", "poly_coeff", "bez_coeff", "poly_time", "x_final", "jerk")
if a.dump:   # child: solve with the library DIRECT_DDP_LIB names and dump
    sys.path.insert(0, ".")
    from direct_b200 import make_batch
    from direct_b200.capi import Solver
    s = Solver(0, "fp64")
    g0, g1 = s.solve_two_stage(make_batch(a.batch, a.knots, a.kind), want_stage0=True)
    np.savez(a.dump, **{f"s{k}_{f}": getattr(g, f) for k, g in enumerate((g0, g1)) for f in FIELDS},
             s0_counts=g0.stats[:, :4], s1_counts=g1.stats[:, :4], kernel_ms=s.stats().kernel_ms)
    s.close()
    sys.exit(0)
out = []
for k, lib in enumerate(a.libs):
    path = f"gpurun_out/_cmp_{k}.npz"
    env = dict(os.environ)
    if lib != "cur":
        env["DIRECT_DDP_LIB"] = lib
    subprocess.check_call([sys.executable, __file__, "--dump", path, "--batch", str(a.batch), "--knots", str(a.knots), "--kind", a.kind], env=env)
    out.append(np.load(path))
same = True
for f in out[0].files:
    if f == "kernel_ms":
        continue
    if not np.array_equal(out[0][f], out[1][f]):
        same = False
        d = np.argwhere(out[0][f] != out[1][f])
        print("DIFFERENT", f, len(d), "entries, first", d[:2].tolist())
print(f"{a.libs[0]} ({float(out[0]['kernel_ms']):.1f} ms) vs {a.libs[1]} ({float(out[1]['kernel_ms']):.1f} ms) on {a.batch} x {a.knots} {a.kind}: "
      f"{'IDENTICAL BITS' if same else 'different'}")
