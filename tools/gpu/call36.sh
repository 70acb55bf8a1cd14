#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --precision fp64 --kind box --max-work 3e7 --reps 2 > gpurun_out/r2y_sweep.md 2> gpurun_out/r2y_sweep.err
cat gpurun_out/r2y_sweep.md; tail -2 gpurun_out/r2y_sweep.err
