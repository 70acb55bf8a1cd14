timeout 600 python -m pytest tests/test_voxel.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/vx_pytest.log; cat gpurun_out/vx_pytest.log
timeout 300 python tools/voxel_report.py --out gpurun_out/voxel_report.json 2>&1 | tee gpurun_out/vx_report.log
timeout 300 python tools/voxel_report.py --itr 2 --map 120 120 30 --pillars 60 2>&1 | tee gpurun_out/vx_report_small.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:convex_test_kernel -c 1 -o gpurun_out/prof_voxel -f python tools/voxel_report.py --reps 1 > gpurun_out/ncu_voxel.log 2>&1
tail -3 gpurun_out/ncu_voxel.log
