set -x
mkdir -p gpurun_out/golden
python tests/golden/make_voxel_golden.py gpurun_out/golden > gpurun_out/vx_golden.log 2>&1
cp gpurun_out/golden/*.npz tests/golden/ 
timeout 600 python -m pytest tests/test_voxel.py -x -q 2>&1 | tail -30 > gpurun_out/vx_pytest.log
timeout 300 python tools/voxel_report.py --out gpurun_out/voxel_report.json > gpurun_out/vx_report.log 2>&1
timeout 300 python tools/voxel_report.py --itr 2 --map 120 120 30 --pillars 60 > gpurun_out/vx_report_small.log 2>&1
tail -3 gpurun_out/vx_golden.log; cat gpurun_out/vx_pytest.log; cat gpurun_out/vx_report.log gpurun_out/vx_report_small.log
