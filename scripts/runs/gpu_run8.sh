#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/run8_tiles.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py -q 2>&1 | tail -3 | tee gpurun_out/run8_tests.log
for t in 1 2 4; do
  echo "== tiles_per_cta=$t" | tee -a gpurun_out/run8_tiles.log
  PIA_ATTN_TILES_PER_CTA=$t timeout 300 python scripts/microbench.py 2>&1 | grep -E "tree_attn|whole" | tee -a gpurun_out/run8_tiles.log
  PIA_ATTN_TILES_PER_CTA=$t timeout 300 python scripts/attn_debug.py 2>&1 | head -15 | tail -13 >> gpurun_out/run8_tiles.log
done
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run8_generate.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/run8_bench_7b.log
