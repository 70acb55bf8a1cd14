#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
DIRECT_DDP_LIB=tools/_variants/lib_trc.so timeout 200 python tools/timeline.py --cycles > gpurun_out/r2v_timeline_cycles.log 2>&1
sed -n 1,2p gpurun_out/r2v_timeline_cycles.log; sed -n 60,103p gpurun_out/r2v_timeline_cycles.log
