#!/bin/bash
mkdir -p gpurun_out
export PIA_PDL=1
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | grep -vE "^\s*$" | tail -40 > gpurun_out/run36_memcheck_kernels.log
echo "rc=$?" >> gpurun_out/run36_memcheck_kernels.log
tail -15 gpurun_out/run36_memcheck_kernels.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_trie.py -q -x -k "golden or differential" 2>&1 | grep -vE "^\s*$" | tail -30 > gpurun_out/run36_memcheck_trie.log
tail -8 gpurun_out/run36_memcheck_trie.log
