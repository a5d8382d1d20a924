#!/bin/bash
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_tree_attention and 1000" 2>&1 | grep -vE "^\s*$" | head -60 | cut -c1-600 > gpurun_out/run38_synccheck_attn.log
head -40 gpurun_out/run38_synccheck_attn.log
timeout 300 python -m pytest tests/test_gpu_trie.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run38_trie.log
