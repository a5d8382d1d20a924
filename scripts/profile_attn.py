# -*- coding: utf-8 -*-
"""k_tree_attn alone, at the two shapes bench.py reports a roofline for, for `ncu --set full`:
    python scripts/profile_attn.py short     n=64 P=384  (the benchmark's mid-generation step), 32 layers of KV
    python scripts/profile_attn.py long      n=64 P=3968, 4 layers of KV
    python scripts/profile_attn.py gqa       n=64 P=384, 32 q heads / 8 kv heads
Numbers printed here are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'short'
dev = torch.device('cuda:0')
if which == 'short':
    r = bench.attention_roofline_long(dev, P=bench.PROMPT_LEN + bench.NEW_TOKENS // 2, layers=32)
elif which == 'gqa':   # Mistral-7B / Mixtral attention geometry (GQA-4) at the benchmark's mid-generation step
    r = bench.attention_roofline_long(dev, P=bench.PROMPT_LEN + bench.NEW_TOKENS // 2, hq=32, hkv=8, layers=32)
elif which == 'gqa_long':   # the same geometry at a 4 k context: where the tensor pipe has work to do (AI 256 FLOP/B)
    r = bench.attention_roofline_long(dev, hq=32, hkv=8, layers=8)
else:
    r = bench.attention_roofline_long(dev)
print({k: r[k] for k in ('shape', 'us_per_launch', 'achieved', 'frac', 'bytes_per_launch')})
