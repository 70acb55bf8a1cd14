#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fixtures" > gpurun_out/r2z_pytest_fixtures.log 2>&1; echo "exit $?" >> gpurun_out/r2z_pytest_fixtures.log
tail -8 gpurun_out/r2z_pytest_fixtures.log
