#!/bin/bash
mkdir -p gpurun_out
for t in 1 2 4; do
  echo "== tiles_per_cta=$t"
  PIA_ATTN_TILES_PER_CTA=$t timeout 300 python scripts/microbench.py 2>&1 | grep -E "tree_attn|whole" | tee -a gpurun_out/run7_tiles.log
done
timeout 300 python scripts/attn_debug.py 2>&1 | head -16 | tee gpurun_out/run7_attn_debug.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k attention 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run7_generate.log
