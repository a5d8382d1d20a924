#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run20_tests.log
for t in 1 2 4; do
  echo "== tiles_per_cta=$t" | tee -a gpurun_out/run20_tiles.log
  PIA_ATTN_TILES_PER_CTA=$t timeout 300 python scripts/microbench.py 2>&1 | grep -E "tree_attn|whole" | tee -a gpurun_out/run20_tiles.log
done
PIA_ATTN_TILES_PER_CTA=1 timeout 300 python scripts/attn_debug.py 2>&1 | head -15 | tee gpurun_out/run20_attn_debug.log
