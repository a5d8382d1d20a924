#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "attn or attention" 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run23_tests.log
timeout 200 python scripts/attn_debug.py 2>&1 | head -15 | tee gpurun_out/run23_attn_debug.log
for pf in 0 "1,0,0,0" "1,0.3,0,0" "1,0.3,0,4000" "1,0.3,0.3,4000" "1,0.5,0.3,4000" "0.5,0,0,3000" "1,0.2,0.2,3000"; do
  PIA_PREFETCH=$pf timeout 200 python scripts/microbench.py --forward-only 2>&1 | grep -E "verify|Error|error" | tee -a gpurun_out/run23_prefetch.log
done
timeout 200 python scripts/microbench.py 2>&1 | grep -E "tree_attn" | tee -a gpurun_out/run23_prefetch.log
