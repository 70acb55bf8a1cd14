#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2l cur > /dev/null 2>&1
O=gpurun_out
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2l_ab.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu >> $O/r2l_ab.log 2>&1
grep -v "cooperation\|small batch\|^rc" $O/r2l_ab.log
