#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/tail_report.py --top 20 > gpurun_out/r2t_tail.log 2>&1
timeout 200 python tools/timeline.py > gpurun_out/r2t_timeline.log 2>&1
cat gpurun_out/r2t_tail.log; tail -40 gpurun_out/r2t_timeline.log
