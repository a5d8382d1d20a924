#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/attn_debug.py 2>&1 | tail -45 | tee gpurun_out/run6_attn_debug.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k attention 2>&1 | tail -5 | tee gpurun_out/run6_kernels.log
timeout 600 python scripts/microbench.py 2>&1 | tail -20 | tee gpurun_out/run6_microbench.log
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run6_generate.log
