#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run22_tests.log
timeout 200 python scripts/microbench.py 2>&1 | tee gpurun_out/run22_micro.log
timeout 200 python scripts/attn_debug.py 2>&1 | head -15 | tee gpurun_out/run22_attn_debug.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r3.csv python scripts/profile_step.py --new 12 --requests 2 > gpurun_out/run22_profile_stdout.log 2>&1
tail -2 gpurun_out/run22_profile_stdout.log
