#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k gemm 2>&1 | tail -12 | tee gpurun_out/run16_tests.log
timeout 900 python scripts/gemm_bench.py 2>&1 | grep -E "cuBLAS|stream-K|tiled" | tee gpurun_out/run16_gemm_bench.log
