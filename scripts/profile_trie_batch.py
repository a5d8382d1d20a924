# -*- coding: utf-8 -*-
"""The batched hier_get scan of bench.py's trie roofline, alone, for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

r = bench.trie_roofline(torch.device('cuda:0'), n_docs=int(os.environ.get('DOCS', 1500)), n_queries=4096)
print({k: r[k] for k in ('achieved', 'frac', 'ms_per_launch', 'bytes_per_launch', 'forest_nodes', 'mean_draft')})
