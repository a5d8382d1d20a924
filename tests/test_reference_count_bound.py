# -*- coding: utf-8 -*-
"""The premise of the pruned query walk (csrc/trie.cu tree_get), checked on the LIVE reference trie: whatever sequence
of put / stream_put / reset_input_freqs / squeeze the reference executes, every node's counts bound its children's
(`_put` adds along root paths, lookahead_cache.py:40-56; `_squeeze` halves top-down and pops whole subtrees, :302-310;
`_reset_input_freq` zeroes top-down, :326-333).  Runs only where the reference checkout exists (the build container);
the CUDA trie asserts the same on its own forests in tests/test_gpu_trie.py."""
import os
import sys

import numpy as np
import pytest

REF = '/root/reference/lookahead'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='the reference checkout only exists in the build container')


def _violations(nodes, parent_freqs, idxs):
    bad = 0
    for node in nodes.values():
        if parent_freqs is not None:
            for k in idxs:
                if node.freqs.get(k, 0.0) > parent_freqs.get(k, 0.0):
                    bad += 1
        bad += _violations(node.children, node.freqs, idxs)
    return bad


@pytest.mark.parametrize('seed,vocab,zipf', [(1, 12, False), (2, 3000, True), (3, 40, False)])
def test_reference_counts_bound_their_children(seed, vocab, zipf):
    sys.path.insert(0, REF)
    try:
        from lookahead.common.lookahead_cache import LookaheadCache
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(seed)

    def toks(k):
        if zipf:
            return np.clip(rng.zipf(1.3, size=k), 3, vocab - 1).tolist()
        return rng.integers(3, vocab, size=k).tolist()

    c = LookaheadCache(eos_ids=[2])
    c.max_node, c.max_output_node = 64, 24          # tiny limits: squeezes actually happen
    checks = 0
    for req in range(1500 if zipf else 300):
        idx = req % 3
        prompt = toks(int(rng.integers(4, 80)))
        c.put(prompt[1:], branch_length=9, final=False, mode='input', idx=idx)
        for _ in range(int(rng.integers(1, 10))):
            c.stream_put(toks(int(rng.integers(1, 9))), branch_length=9, final=False, mode='output', idx=idx)
        c.stream_put([], branch_length=9, final=True, mode='output', idx=idx)   # reset_input_freqs + squeeze (:403-406)
        if req % 50 == 49:
            for tree in c.mem.values():
                assert _violations(tree.nodes, None, (-1, 0, 1, 2)) == 0
                checks += 1
    assert checks > 20
