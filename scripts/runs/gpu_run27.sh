#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r4_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/run27_launches_stdout.log 2>&1
tail -1 gpurun_out/run27_launches_stdout.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tree_attn -s 40 -c 2 -o gpurun_out/prof_tree_attn_r2 -f python scripts/profile_step.py --new 8 --requests 1 > gpurun_out/run27_ncu_attn.log 2>&1
tail -1 gpurun_out/run27_ncu_attn.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -s 40 -c 2 -o gpurun_out/prof_gemm_ws_r2 -f python scripts/profile_step.py --new 8 --requests 1 > gpurun_out/run27_ncu_gemm.log 2>&1
tail -1 gpurun_out/run27_ncu_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_get -s 1 -c 1 -o gpurun_out/prof_trie_batch_r2 -f python scripts/profile_trie_batch.py > gpurun_out/run27_ncu_trie.log 2>&1
tail -1 gpurun_out/run27_ncu_trie.log
ls -la gpurun_out/*.ncu-rep
