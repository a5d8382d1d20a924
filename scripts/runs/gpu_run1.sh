#!/bin/bash
# first GPU contact: trie parity tests (with compute-sanitizer on a short subset)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_trie.py -x -q 2>&1 | tail -40 | tee gpurun_out/run1_pytest.log
