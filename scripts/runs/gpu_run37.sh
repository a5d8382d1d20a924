#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_trie.py -q -x -k "golden or differential" 2>&1 | grep -vE "^\s*$" | tail -40 > gpurun_out/run37_racecheck_trie.log
tail -12 gpurun_out/run37_racecheck_trie.log
timeout 500 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trie.py -q -x 2>&1 | grep -vE "^\s*$" | tail -30 > gpurun_out/run37_synccheck.log
tail -8 gpurun_out/run37_synccheck.log
