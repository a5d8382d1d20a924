#!/bin/bash
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_tree_attention" 2>&1 | grep -vE "^\s*$" | grep -E "=========|passed|failed" | head -40 | cut -c1-900 > gpurun_out/run40_synccheck.log
cat gpurun_out/run40_synccheck.log
