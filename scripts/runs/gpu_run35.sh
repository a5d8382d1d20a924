#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "grouped or gemm" 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run35_tests.log
timeout 400 python -m pytest tests/test_gpu_generate.py -q -x 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee -a gpurun_out/run35_tests.log
( time timeout 500 python bench.py --model mixtral-8x7b --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/run35_mixtral.log 2>&1
grep '^{' gpurun_out/run35_mixtral.log | tail -1 | cut -c1-800
tail -5 gpurun_out/run35_mixtral.log | cut -c1-300
