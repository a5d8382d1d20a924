#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 --model tiny 2>&1 | tail -2 | tee gpurun_out/run9_bench_tiny_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 4 --warmup 3 2>&1 | tail -2 | tee gpurun_out/run9_bench_7b_n2.log
