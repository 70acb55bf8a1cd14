#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k multi_device > $O/r2z_pytest_multi.log 2>&1; echo "exit $?" >> $O/r2z_pytest_multi.log
tail -3 $O/r2z_pytest_multi.log
timeout 200 python tools/multi_gpu_capi.py --total 8192 --gpus 1 2 --reps 2 2>&1 | tail -3
