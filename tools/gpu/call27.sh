#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/stress_tail.py --reps 3 > gpurun_out/r2v_stress.log 2>&1; echo "rc $?" >> gpurun_out/r2v_stress.log
tail -70 gpurun_out/r2v_stress.log
