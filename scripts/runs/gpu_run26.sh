#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run26_tests.log
nproc | tee gpurun_out/run26_nproc.log
( time timeout 420 python bench.py --steps 8 --warmup 3 ) > gpurun_out/run26_bench.log 2>&1
tail -5 gpurun_out/run26_bench.log | cut -c1-1500
( time timeout 400 python bench.py --impl reference --steps 8 --warmup 3 ) > gpurun_out/run26_ref.log 2>&1
tail -5 gpurun_out/run26_ref.log | cut -c1-1200
