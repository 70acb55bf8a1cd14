#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_balancing or full_size" > gpurun_out/r2u_pytest_tail.log 2>&1; echo "exit $?" >> gpurun_out/r2u_pytest_tail.log
tail -15 gpurun_out/r2u_pytest_tail.log
