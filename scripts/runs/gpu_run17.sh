#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py tests/test_gpu_generate.py -q 2>&1 | tail -6 | tee gpurun_out/run17_tests.log
timeout 600 python scripts/microbench.py 2>&1 | grep -E "whole|tree_attn" | tee gpurun_out/run17_micro.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/run17_bench_7b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_get -s 1 -c 1 -o gpurun_out/prof_trie_batch_r1 -f python scripts/profile_trie_batch.py > gpurun_out/run17_ncu_trie.log 2>&1
tail -2 gpurun_out/run17_ncu_trie.log
