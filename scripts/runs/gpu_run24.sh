#!/bin/bash
mkdir -p gpurun_out
for pf in 0 "1,0,0,0" "1,0.3,0,0" "1,0.3,0,6000" "1,0.3,0.3,6000" "1,0.5,0.3,0" "1,0.2,0.2,4000" "0,0,0.3,0"; do
  PIA_PREFETCH=$pf timeout 200 python scripts/microbench.py --forward-only 2>&1 | grep -E "verify|Error|error" | tee -a gpurun_out/run24_prefetch.log
done
