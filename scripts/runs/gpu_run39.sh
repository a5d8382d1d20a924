#!/bin/bash
mkdir -p gpurun_out
for k in "2-2-0-64-0" "4-2-37-33-0" "32-32-300-64-0"; do
  echo "=== $k" >> gpurun_out/run39_synccheck.log
  timeout 100 compute-sanitizer --tool synccheck --print-limit 4 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_tree_attention and $k" 2>&1 | grep -vE "^\s*$" | grep -E "=========|passed|failed" | head -24 | cut -c1-700 >> gpurun_out/run39_synccheck.log
done
cat gpurun_out/run39_synccheck.log | head -80
