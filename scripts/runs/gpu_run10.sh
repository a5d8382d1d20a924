#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or partials" 2>&1 | tail -15 | tee gpurun_out/run10_gemm_tests.log
timeout 900 python scripts/gemm_bench.py 2>&1 | tail -30 | tee gpurun_out/run10_gemm_bench.log
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run10_generate.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/run10_smoke.log
