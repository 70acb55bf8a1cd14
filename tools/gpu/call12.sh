#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 120 python tools/timeline.py --first 547 > $O/r2n_timeline.log 2>&1
timeout 120 python tools/cycle_report.py --batch 4096 --tag cur >> $O/r2n_timeline.log 2>&1
awk 'NR<=3 || NR%6==0' $O/r2n_timeline.log | head -40; tail -5 $O/r2n_timeline.log
