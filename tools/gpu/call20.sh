#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2u_pytest_gpu.log
tail -3 gpurun_out/r2u_pytest_gpu.log
bash tools/gpu/sanitize.sh memcheck synccheck
timeout 200 python tools/latency_b1.py > gpurun_out/r2u_latency_b1.log 2>&1; tail -6 gpurun_out/r2u_latency_b1.log | head -5
