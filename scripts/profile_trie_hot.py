# -*- coding: utf-8 -*-
"""One hier_get over the hot subtree of a production-size forest (query (3, 3) on 1500 phrase-bank documents: ~137 k
nodes below the match) for `ncu --set full --import-source on -k regex:k_get`: where does a single hot query spend its
0.8 ms?  Numbers printed here are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache  # noqa: E402

dev = torch.device('cuda:0')
big = LookaheadCache(eos_ids=[2], device=dev, vocab_capacity=32000, node_capacity=1 << 23)
for d in bench.phrase_bank_prompts(1500, 32000, seed=7):
    big.put(d, branch_length=9, mode='output', idx=-1)
seq = torch.zeros((1, 64), dtype=torch.int32, device=dev)
seq[0, :2] = torch.tensor([3, 3], dtype=torch.int32, device=dev)
n = torch.tensor([2], dtype=torch.int32, device=dev)
prof = torch.cuda.cudart()
for rep in range(3):
    if rep == 2:
        prof.cudaProfilerStart()
    o = big.get_device(seq, n, 64, 8, min_output_size=32)
    torch.cuda.synchronize()
prof.cudaProfilerStop()
print('draft n', int(o['n'][0]), big.stats())
