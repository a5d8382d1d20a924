#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee gpurun_out/run32_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/run32_smoke.log
( time timeout 500 python bench.py --steps 8 --warmup 3 ) > gpurun_out/run32_bench.log 2>&1
tail -4 gpurun_out/run32_bench.log | cut -c1-600
( time timeout 500 python bench.py --impl reference --steps 8 --warmup 3 ) > gpurun_out/run32_ref.log 2>&1
tail -4 gpurun_out/run32_ref.log | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r5_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/run32_launches_stdout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tree_attn -s 40 -c 2 -o gpurun_out/prof_attn_short_r3 -f python scripts/profile_attn.py short > gpurun_out/run32_ncu_attn_short.log 2>&1
tail -1 gpurun_out/run32_ncu_attn_short.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tree_attn -s 10 -c 2 -o gpurun_out/prof_attn_long_r3 -f python scripts/profile_attn.py long > gpurun_out/run32_ncu_attn_long.log 2>&1
tail -1 gpurun_out/run32_ncu_attn_long.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -s 40 -c 4 -o gpurun_out/prof_gemm_ws_r3 -f python scripts/profile_step.py --new 8 --requests 1 > gpurun_out/run32_ncu_gemm.log 2>&1
tail -1 gpurun_out/run32_ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_get -s 1 -c 1 -o gpurun_out/prof_trie_batch_r3 -f python scripts/profile_trie_batch.py > gpurun_out/run32_ncu_trie.log 2>&1
tail -1 gpurun_out/run32_ncu_trie.log
