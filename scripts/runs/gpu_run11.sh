#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py -q 2>&1 | tail -8 | tee gpurun_out/run11_tests.log
timeout 900 python scripts/gemm_bench.py 2>&1 | tail -40 | tee gpurun_out/run11_gemm_bench.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/run11_bench_7b.log
