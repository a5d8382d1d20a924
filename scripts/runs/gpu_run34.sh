#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python bench.py --model mistral-7b --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/run34_mistral.log 2>&1
grep '^{' gpurun_out/run34_mistral.log | tail -1 | cut -c1-900
tail -4 gpurun_out/run34_mistral.log | grep real
( time timeout 600 python bench.py --model mixtral-8x7b --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/run34_mixtral.log 2>&1
grep '^{' gpurun_out/run34_mixtral.log | tail -1 | cut -c1-900
tail -6 gpurun_out/run34_mixtral.log | cut -c1-300
