#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_gpu_trie.py tests/test_gpu_kernels.py tests/test_gpu_generate.py; do
  echo "=== $t"
  timeout 600 python -m pytest $t -x -q 2>&1 | tail -45 | tee gpurun_out/run2_$(basename $t .py).log
done
