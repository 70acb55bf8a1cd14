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
for B, first in ((1, 7), (1, 547), (1, 1137), (4, 545), (16, 540), (64, 500), (256, 400)):
    pb = make_batch(B, 100, "box", first=first)
    row = []
    for g in (1, 64, 148, 222, 296):
        os.environ["DIRECT_DDP_MIN_GRID"] = str(g)
        best = 1e9
        for _ in range(4):
            s.solve_two_stage(pb)
            best = min(best, s.stats().kernel_ms)
        row.append(f"{g}: {best:7.2f}")
    print(f"B {B:4d} first {first:5d} kernel ms by min grid  " + "  ".join(row), flush=True)
s.close()
PY
timeout 600 python /tmp/mg.py > gpurun_out/r2w_mingrid2.log 2>&1
cat gpurun_out/r2w_mingrid2.log
