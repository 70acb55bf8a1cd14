#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
DIRECT_DDP_LIB=tools/_variants/lib_trc.so timeout 200 python tools/tail_who.py > gpurun_out/r2y_tail_who.log 2>&1
cat gpurun_out/r2y_tail_who.log
