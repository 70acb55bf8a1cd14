#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cat > /tmp/mg.py <<'PY'
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from direct_b200 import make_batch
from direct_b200.capi import Solver
s = Solver(0, "fp64")
for B, first in ((700, 0), (1024, 0), (1100, 0)):
    pb = make_batch(B, 100, "box", first=first)
    row = []
    for g in ("", "296"):
        if g: os.environ["DIRECT_DDP_MIN_GRID"] = g
        else: os.environ.pop("DIRECT_DDP_MIN_GRID", None)
        best = 1e9
        for _ in range(4):
            s.solve_two_stage(pb)
            best = min(best, s.stats().kernel_ms)
        row.append(f"{g or 'default'}: {best:7.2f} (grid {s.stats().grid_blocks})")
    print(f"B {B:4d} kernel ms by min grid  " + "  ".join(row), flush=True)
s.close()
PY
timeout 600 python /tmp/mg.py > gpurun_out/r2w_mingrid3.log 2>&1
cat gpurun_out/r2w_mingrid3.log
timeout 300 python tools/latency_b1.py --reps 5 > gpurun_out/r2w_latency_b1.log 2>&1; grep "^|" gpurun_out/r2w_latency_b1.log
timeout 300 python tools/latency_b1.py --reps 3 --knots 100 --first 547 > gpurun_out/r2w_latency_b1_hard.log 2>&1; grep "^|" gpurun_out/r2w_latency_b1_hard.log
