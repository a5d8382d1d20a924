#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trie.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -20 | tee gpurun_out/run28_tests.log
timeout 300 python scripts/profile_trie_batch.py 2>&1 | tail -2 | tee gpurun_out/run28_trie_batch.log
timeout 300 python scripts/microbench.py 2>&1 | grep -E "trie|draft" | tee gpurun_out/run28_micro.log
