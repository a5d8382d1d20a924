#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/microbench.py 2>&1 | tail -20 | tee gpurun_out/run5_microbench.log
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run5_generate.log
