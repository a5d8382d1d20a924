#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run31_tests.log
for gs in "gate_up" "gate_up,o" "gate_up,down" "gate_up,qkv2" "gate_up,o,down" "gate_up,qkv2,o,down"; do
  echo "== PIA_GEMM_SET=$gs" | tee -a gpurun_out/run31_gemmset.log
  PIA_GEMM_SET=$gs timeout 200 python scripts/microbench.py --forward-only 2>&1 | grep -E "verify|Error|error" | tee -a gpurun_out/run31_gemmset.log
done
