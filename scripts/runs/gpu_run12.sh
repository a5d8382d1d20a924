#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generate.py tests/test_gpu_kernels.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run12_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/run12_smoke.log
timeout 900 python scripts/gemm_bench.py 2>&1 | tail -40 | tee gpurun_out/run12_gemm_bench.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/run12_bench_7b.log
PIA_GEMM=0 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/run12_bench_7b_cublas.log
