#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py tests/test_gpu_generate.py -q 2>&1 | tail -4 | tee gpurun_out/run19_tests.log
timeout 600 python scripts/microbench.py 2>&1 | grep -E "whole|tree_attn|trie|draft" | tee gpurun_out/run19_micro.log
( time timeout 1500 python bench.py --steps 8 --warmup 3 ) > gpurun_out/run19_bench_full.log 2>&1
tail -c 600 gpurun_out/run19_bench_full.log
( time timeout 1500 python bench.py --impl reference --steps 8 --warmup 3 ) > gpurun_out/run19_bench_reference.log 2>&1
tail -c 900 gpurun_out/run19_bench_reference.log
