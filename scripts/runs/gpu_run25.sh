#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run25_tests.log
for gs in "gate_up" "gate_up,qkv" "gate_up,down" "gate_up,o" "gate_up,qkv,down" "gate_up,qkv,o,down"; do
  echo "== PIA_GEMM_SET=$gs" | tee -a gpurun_out/run25_gemmset.log
  PIA_GEMM_SET=$gs timeout 200 python scripts/microbench.py --forward-only 2>&1 | grep -E "verify|Error|error" | tee -a gpurun_out/run25_gemmset.log
done
