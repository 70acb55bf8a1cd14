#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2i cur > /dev/null 2>&1
O=gpurun_out
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2i_ab.log 2>&1
timeout 120 python tools/spec_check.py >> $O/r2i_ab.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_replay.py -x -q -m gpu >> $O/r2i_ab.log 2>&1
grep -v "cooperation\|smoke\|small batch\|^rc" $O/r2i_ab.log
