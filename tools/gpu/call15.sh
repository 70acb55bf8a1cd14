#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voxel.py -x -q -m gpu > gpurun_out/r2t_voxel_test.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2t_voxel_test.log
timeout 600 python tools/voxel_report.py --out gpurun_out/r2t_voxel_report.json > gpurun_out/r2t_voxel_report.log 2>&1; echo "report rc $?" >> gpurun_out/r2t_voxel_report.log
tail -15 gpurun_out/r2t_voxel_test.log; tail -20 gpurun_out/r2t_voxel_report.log
