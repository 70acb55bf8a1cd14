#!/bin/bash
# single-GPU record run of the round: full gpu test suite, smoke, bench (ours + reference arm), configs[3], B=1 latency, sweeps
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
T=${1:-r2p}
nvidia-smi -L > $O/${T}_host.txt; nproc >> $O/${T}_host.txt; lscpu | grep "Model name" >> $O/${T}_host.txt
timeout 500 python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${T}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "exit $?" >> $O/${T}_smoke.log
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 300 python bench.py --kind poly --knots 200 --no-model-b --cpu-sample 512 --screen-sample 0 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
timeout 300 python bench.py --batch 65536 --steps 3 --no-model-b --no-cpu-baseline > $O/${T}_bench_64k.json 2> $O/${T}_bench_64k.err
timeout 300 python tools/latency_b1.py > $O/${T}_latency_b1.log 2>&1
tail -3 $O/${T}_pytest_gpu.log; tail -4 $O/${T}_smoke.log; head -c 400 $O/${T}_bench.json; echo; head -c 300 $O/${T}_bench_ref.json; echo; head -c 300 $O/${T}_bench_cfg3.json; echo; head -c 300 $O/${T}_bench_64k.json; echo; tail -8 $O/${T}_latency_b1.log
