#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
DIRECT_DDP_LIB=tools/_variants/lib_trc.so timeout 200 python tools/timeline.py --cycles --batch 1 --first 547 > gpurun_out/r2z_timeline_b1.log 2>&1
sed -n 1,2p gpurun_out/r2z_timeline_b1.log; sed -n 40,70p gpurun_out/r2z_timeline_b1.log
