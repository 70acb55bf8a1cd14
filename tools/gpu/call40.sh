#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python tools/stress_tail.py --reps 3 > gpurun_out/r2z_stress.log 2>&1; echo "rc $?" >> gpurun_out/r2z_stress.log; tail -3 gpurun_out/r2z_stress.log
bash tools/gpu/sanitize.sh memcheck synccheck initcheck
