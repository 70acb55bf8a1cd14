#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ipddp -s 1 -c 1 -o gpurun_out/prof_r2w_b1 -f python tools/profile_one.py --batch 1 --first 547 > gpurun_out/ncu_r2w_b1.log 2>&1
tail -3 gpurun_out/ncu_r2w_b1.log
