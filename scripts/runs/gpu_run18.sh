#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gemm_bench.py 2>&1 | grep -E "silu|rmsnorm" | tee gpurun_out/run18_gemm_bench.log
