#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/run14_bench_7b.log
timeout 900 python scripts/gemm_bench.py 2>&1 | grep -E "gate_up|qkv" | tee gpurun_out/run14_gemm_bench.log
timeout 600 python -m pytest tests/test_gpu_generate.py tests/test_gpu_kernels.py tests/test_gpu_trie.py -q 2>&1 | tail -3 | tee gpurun_out/run14_tests.log
