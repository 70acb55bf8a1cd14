#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2q_host2.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k multi_device > $O/r2q_pytest_multi.log 2>&1; echo "exit $?" >> $O/r2q_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2q_bench_2gpu.json 2> $O/r2q_bench_2gpu.err
timeout 300 python tools/multi_gpu_capi.py --total 8192 --gpus 1 2 --reps 2 > $O/r2q_capi_strong2.jsonl 2> $O/r2q_capi_strong2.err
tail -3 $O/r2q_pytest_multi.log; head -c 500 $O/r2q_bench_2gpu.json; echo; tail -3 $O/r2q_bench_2gpu.err; cat $O/r2q_capi_strong2.jsonl; tail -3 $O/r2q_capi_strong2.err
