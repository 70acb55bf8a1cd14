#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2r_host8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2r_bench_8gpu.json 2> $O/r2r_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --batch 8192 --steps 3 --warmup 3 > $O/r2r_bench_8gpu_cfg2.json 2> $O/r2r_bench_8gpu_cfg2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2r_bench_4gpu.json 2> $O/r2r_bench_4gpu.err
timeout 400 python tools/multi_gpu_capi.py --total 65536 --gpus 8 4 --reps 2 > $O/r2r_capi_strong.jsonl 2> $O/r2r_capi_strong.err
head -c 300 $O/r2r_bench_8gpu.json; echo; head -c 300 $O/r2r_bench_8gpu_cfg2.json; echo; head -c 300 $O/r2r_bench_4gpu.json; echo; cat $O/r2r_capi_strong.jsonl; tail -2 $O/r2r_capi_strong.err
