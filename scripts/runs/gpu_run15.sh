#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
  echo "== PIA_PDL=$pdl" | tee -a gpurun_out/run15.log
  PIA_PDL=$pdl timeout 600 python -m pytest tests/test_gpu_generate.py tests/test_gpu_kernels.py -q 2>&1 | tail -3 | tee -a gpurun_out/run15.log
  PIA_PDL=$pdl timeout 600 python scripts/microbench.py 2>&1 | grep -E "whole|tree_attn|rmsnorm|rope|silu" | tee -a gpurun_out/run15.log
  PIA_PDL=$pdl timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/run15_bench_pdl$pdl.log
done
