#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_generate.jsonl
timeout 900 python -m pytest tests/test_gpu_generate.py -q 2>&1 | grep -E "^(FAILED|E  |[0-9]+ (passed|failed))|passed|failed" | head -40 | tee gpurun_out/run4_generate.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/run4_smoke.log
# launch list of a short 7B run (device time per launch; shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv python scripts/profile_step.py --new 24 --requests 2 > gpurun_out/run4_profile_stdout.log 2>&1
tail -3 gpurun_out/run4_profile_stdout.log
# full capture of the attention kernel and of the trie get kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tree_attn -s 40 -c 2 -o gpurun_out/prof_tree_attn_r1 -f python scripts/profile_step.py --new 8 --requests 1 > gpurun_out/run4_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_get -s 2 -c 2 -o gpurun_out/prof_trie_get_r1 -f python scripts/profile_step.py --new 16 --requests 2 > gpurun_out/run4_ncu_get.log 2>&1
ls -la gpurun_out | tail -12
