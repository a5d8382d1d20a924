#!/bin/bash
mkdir -p gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/run33_bench2.log 2>&1
grep '^{' gpurun_out/run33_bench2.log | tail -1 | cut -c1-700
tail -4 gpurun_out/run33_bench2.log | grep real
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/run33_ref2.log 2>&1
grep '^{' gpurun_out/run33_ref2.log | tail -1 | cut -c1-300
tail -4 gpurun_out/run33_ref2.log | grep real
