#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/lpt_probe.py > gpurun_out/r2t_lpt.log 2>&1
cat gpurun_out/r2t_lpt.log
