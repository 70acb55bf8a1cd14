#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest_gpu_final.log 2>&1; echo "exit $?" >> gpurun_out/r2z_pytest_gpu_final.log
tail -4 gpurun_out/r2z_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
