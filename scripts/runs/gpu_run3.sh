#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | tail -60 | tee gpurun_out/run3_generate.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/run3_smoke.log
timeout 900 python bench.py --model tiny --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/run3_bench_tiny.log
timeout 1500 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/run3_bench_7b.log
