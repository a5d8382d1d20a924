#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/microbench.py 2>&1 | grep -E "whole|tree_attn" | tee gpurun_out/run13_micro.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python scripts/profile_step.py --new 12 --requests 2 > gpurun_out/run13_profile_stdout.log 2>&1
tail -2 gpurun_out/run13_profile_stdout.log
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -10 | tee gpurun_out/run13_tests.log
