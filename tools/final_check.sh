# Round-end verification on one B200: full GPU suite, smoke, bench (both arms), model (B) report lines; with NCU=1 also the launch list of the
# bench command and the ncu --set full captures of the pair kernel (profiles/r1j, r1k).
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cat gpurun_out/bench_final_ref.json
[ -n "$NCU" ] && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_r1j.log 2>&1
[ -n "$NCU" ] && tail -2 gpurun_out/b_ncu_r1j.log
[ -n "$NCU" ] && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gddp_pair -c 1 -o gpurun_out/prof_gddp_pair -f python tools/gddp_report.py --reps 1 > gpurun_out/ncu_gddp_pair.log 2>&1
[ -n "$NCU" ] && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gddp_pair -c 1 -o gpurun_out/prof_gddp_pair64k -f python tools/gddp_report.py --reps 1 --batch 65536 > gpurun_out/ncu_gddp_pair64k.log 2>&1
timeout 200 python tools/gddp_report.py | tee gpurun_out/gddp_final.log
timeout 200 python tools/gddp_report.py --batch 16384 | tee -a gpurun_out/gddp_final.log
timeout 200 python tools/gddp_report.py --batch 65536 --reps 3 | tee -a gpurun_out/gddp_final.log
timeout 200 python tools/gddp_report.py --precision fp64 | tee -a gpurun_out/gddp_final.log
timeout 200 python tools/gddp_report.py --knots 200 | tee -a gpurun_out/gddp_final.log
timeout 200 python tools/gddp_report.py --model dint --knots 50 | tee -a gpurun_out/gddp_final.log
