# -*- coding: utf-8 -*-
"""Device time of single hier_get calls (CUDA events, warm): a hot query (the 137 k-node subtree below (3, 3) of a
1500-document phrase-bank forest), a cold one, and the batched scan.  PIA_TRIE_GET_CLUSTER selects CTAs per row."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache  # noqa: E402

dev = torch.device('cuda:0')
big = LookaheadCache(eos_ids=[2], device=dev, vocab_capacity=32000, node_capacity=1 << 23)
docs = bench.phrase_bank_prompts(1500, 32000, seed=7)
for d in docs:
    big.put(d, branch_length=9, mode='output', idx=-1)


REPS = int(os.environ.get('REPS', 20))


def timed(seq, n, reps=REPS, batch=1):
    for _ in range(3):
        o = big.get_device(seq, n, 64, 8, min_output_size=32, batch=batch) if batch > 1 else big.get_device(seq, n, 64, 8, min_output_size=32)
    torch.cuda.synchronize()
    st0 = big.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        o = big.get_device(seq, n, 64, 8, min_output_size=32, batch=batch) if batch > 1 else big.get_device(seq, n, 64, 8, min_output_size=32)
    e1.record()
    torch.cuda.synchronize()
    st1 = big.stats()
    nv = (st1['nodes_visited'] - st0['nodes_visited']) / reps
    return e0.elapsed_time(e1) / reps * 1e3, nv, o


seq = torch.zeros((1, 64), dtype=torch.int32, device=dev)
seq[0, :2] = torch.tensor([3, 3], dtype=torch.int32, device=dev)
n = torch.tensor([2], dtype=torch.int32, device=dev)
us, nv, o = timed(seq, n)
print('prune', os.environ.get('PIA_TRIE_PRUNE', '1'), 'cluster', os.environ.get('PIA_TRIE_GET_CLUSTER', '8'), 'hot query: %.1f us, %d nodes visited, draft n %d, ids %s'
      % (us, nv, int(o['n'][0]), o['ids'][0, :8].tolist()))
doc = docs[17]
seq2 = torch.zeros((1, 64), dtype=torch.int32, device=dev)
seq2[0, :2] = torch.tensor(doc[40:42], dtype=torch.int32, device=dev)
us, nv, o = timed(seq2, n)
print('cold query: %.1f us, %d nodes visited, draft n %d' % (us, nv, int(o['n'][0])))
# a sample of ordinary queries (adjacent pairs of the documents), each timed alone
import numpy as np  # noqa: E402
rng = np.random.default_rng(5)
res = []
for _ in range(int(os.environ.get('SAMPLE', 150))):
    d = docs[int(rng.integers(0, len(docs)))]
    j = int(rng.integers(0, len(d) - 2))
    sq = torch.zeros((1, 64), dtype=torch.int32, device=dev)
    sq[0, :2] = torch.tensor(d[j:j + 2], dtype=torch.int32, device=dev)
    us, nv, o = timed(sq, n, reps=5)
    res.append((us, nv, int(o['n'][0])))
a = np.array(res)
order = np.argsort(a[:, 1])
print('sample of %d pair queries: mean %.1f us, p50 %.1f, p90 %.1f, max %.1f; nodes visited p50 %d p90 %d max %d'
      % (len(a), a[:, 0].mean(), np.percentile(a[:, 0], 50), np.percentile(a[:, 0], 90), a[:, 0].max(),
         np.percentile(a[:, 1], 50), np.percentile(a[:, 1], 90), a[:, 1].max()))
for lo in range(0, len(a), len(a) // 5):
    sl = a[order[lo:lo + len(a) // 5]]
    print('   visited %6d..%6d: mean %.1f us (draft n %.0f)' % (sl[:, 1].min(), sl[:, 1].max(), sl[:, 0].mean(), sl[:, 2].mean()))
if os.environ.get('SKIP_SCAN'):
    sys.exit(0)
r = bench.trie_roofline(dev, n_docs=1500, n_queries=4096)
print({k: r[k] for k in ('achieved', 'frac', 'ms_per_launch', 'us_per_get', 'bytes_per_launch', 'bytes_visited_per_launch',
                          'achieved_on_visited_bytes', 'pruned_walk_equals_full_walk', 'forest_nodes', 'mean_draft')})
