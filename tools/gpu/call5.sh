#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/r2f_pytest_all.log 2>&1; echo "all exit $?" >> $O/r2f_pytest_all.log
tail -4 $O/r2f_pytest_all.log
bash tools/gpu/ncu_one.sh r2f
