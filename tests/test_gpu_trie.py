# -*- coding: utf-8 -*-
"""Parity of the CUDA trie (csrc/trie.cu through the C ABI) with the reference:
 * the reference's own unit vectors (lookahead/tests/test_lookahead_cache.py:16-45) through the Tree mirror,
 * op streams recorded from the live reference (tests/golden/*.json),
 * seeded differential streams against the CPU oracle at sizes the fixtures cannot hold."""
import numpy as np
import pytest

from tests import replay as R

pytestmark = pytest.mark.gpu

STREAMS = ['trie_survey_a1.json', 'trie_small_v12.json', 'trie_small_v6_stop.json', 'trie_small_v30_batch.json',
           'trie_small_v8_dl128.json', 'trie_squeeze.json', 'trie_zipf.json']


def _gpu_cache_cls():
    from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache
    return LookaheadCache


def _stats(c):
    s = c.stats()
    assert s['error_flags'] == 0, s
    return {'n_trees': s['n_trees'], 'n_update_trees': s['n_update_trees']}


def test_reference_unit_vectors_literal():
    from painlessinferenceacceleration_b200.common.lookahead_cache import Tree
    tree = Tree(1)
    tree.put([1, 2, 3, 4], mode='output', idx=-1)
    ids, mask, sizes = tree.get([1], max_size=63, max_length=8, min_input_size=0, min_output_size=0,
                                output_weight=1e-4, mode='mix', idx=0)
    assert ids == [1, 2, 3, 4], ids
    assert mask.shape == (4, 4)
    assert np.sum(np.abs(mask - np.array([[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1]]))) == 0, mask
    tree = Tree(1)
    tree.put([1, 2, 3], mode='output', idx=-1)
    tree.put([1, 2, 4], mode='output', idx=-1)
    ids, mask, sizes = tree.get([1], max_size=63, max_length=8, min_input_size=0, min_output_size=0,
                                output_weight=1e-4, mode='mix', idx=0)
    assert ids == [1, 2, 3, 4], ids
    assert np.sum(np.abs(mask - np.array([[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 0, 1]]))) == 0, mask
    assert tree.n_node == 4 and tree.n_output_node == 4


@pytest.mark.parametrize('name', STREAMS)
def test_recorded_streams(name):
    fx = R.load(name)
    cache = R.make_cache(_gpu_cache_cls(), fx['ctor'])
    n = R.replay(cache, fx['ops'], stats_fn=_stats, tag=name)
    assert n > 0
    assert cache.stats()['error_flags'] == 0


def _assert_counts_bound_children(cache):
    """the invariant the pruned query walk relies on (csrc/trie.cu tree_get; lookahead_cache.py:40-56, :302-310,
    :326-333): after any sequence of put / stream_put / squeeze / reset, every node's fo and fi are >= those of each of
    its children (a tree's root record carries no counts of its own and is exempt)"""
    nodes, edges, root_of, _n, _o = cache.snapshot()
    nch, child, cap, fo, fi = nodes['n_child'], nodes['child'], nodes['cap'], nodes['fo'], nodes['fi']
    roots = set(int(r) for r in root_of[root_of >= 0])
    stack = list(roots)
    checked = 0
    while stack:
        i = stack.pop()
        if nch[i] == 0:
            continue
        kids = [int(child[i])] if cap[i] == 0 else edges[child[i]:child[i] + nch[i], 1].tolist()
        for c in kids:
            if i not in roots:
                assert fo[c] <= fo[i] and fi[c] <= fi[i], (i, c, fo[i], fo[c], fi[i], fi[c])
                checked += 1
            stack.append(c)
    assert checked > 0


def _random_stream(seed, V, n_req, stop=(), dl=64, bl=8, zipf=False):
    """yields op tuples; the same generator drives oracle and GPU"""
    rng = np.random.default_rng(seed)

    def toks(k):
        if zipf:
            return np.clip(rng.zipf(1.3, size=k), 3, V - 1).tolist()
        return rng.integers(0, V, size=k).tolist()

    ops = []
    for req in range(n_req):
        prompt = toks(int(rng.integers(2, 200)))
        ops.append(('put', prompt[1:], bl + 1, False, 'input', 0))
        seq = list(prompt)
        for _ in range(int(rng.integers(1, 30))):
            ops.append(('hier_get', seq[-2:], dl, bl, 0, dl // 2, 'mix', 0))
            if rng.random() < 0.1:
                ops.append(('hier_get', seq[-2:], dl, bl, 0, 0, str(rng.choice(['input', 'output'])), 0))
            if rng.random() < 0.1:
                ops.append(('one_get', seq[-2:], dl, bl, 0, 0, 'mix', 0))
            new = toks(int(rng.integers(1, bl + 2)))
            seq.extend(new)
            ops.append(('stream_put', new, bl + 1, False, 'output', 0))
        ops.append(('stream_put', [], bl + 1, True, 'output', 0))
        if req % 5 == 0:
            ops.append(('put', toks(int(rng.integers(2, 300))), bl + 1, False, 'output', -1))
    return ops


@pytest.mark.parametrize('seed,V,n_req,zipf', [(11, 9, 60, False), (12, 300, 60, False), (13, 32000, 80, True),
                                               (14, 5, 40, False)])
def test_differential_vs_oracle(seed, V, n_req, zipf):
    from oracle.trie import OracleLookaheadCache
    gpu = _gpu_cache_cls()(eos_ids=[2])
    cpu = OracleLookaheadCache(eos_ids=[2])
    checked = 0
    for k, op in enumerate(_random_stream(seed, V, n_req, zipf=zipf)):
        name = op[0]
        if name in ('put', 'stream_put'):
            _, ids, b, final, mode, idx = op
            for c in (cpu, gpu):
                getattr(c, name)(list(ids), branch_length=b, final=final, mode=mode, idx=idx)
        else:
            _, q, dl, b, mi, mo, mode, idx = op
            outs = []
            for c in (cpu, gpu):
                ids, m, sizes = getattr(c, name)(list(q), decoding_length=dl, branch_length=b, min_input_size=mi,
                                                 min_output_size=mo, mode=mode, idx=idx)
                outs.append(([int(x) for x in ids], R.mask_rows(m), [int(x) for x in sizes]))
            assert outs[0] == outs[1], f'op#{k} {op}\n cpu={outs[0]}\n gpu={outs[1]}'
            checked += 1
    assert checked > 50
    s = gpu.stats()
    assert s['error_flags'] == 0 and s['n_trees'] == cpu.n_trees()
    _assert_counts_bound_children(gpu)


def test_batched_get_matches_single():
    """4096 concurrent hier_get rows in one launch give the same rows as one-at-a-time calls (the roofline
    benchmark's batched scan uses this path)"""
    LookaheadCache = _gpu_cache_cls()
    rng = np.random.default_rng(5)
    c = LookaheadCache(eos_ids=[2])
    docs = [np.clip(rng.zipf(1.3, size=256), 3, 31999).tolist() for _ in range(64)]
    for d in docs:
        c.put(d, branch_length=9, mode='output', idx=-1)
    qs = []
    for _ in range(512):
        d = docs[int(rng.integers(0, len(docs)))]
        j = int(rng.integers(0, len(d) - 2))
        qs.append(d[j:j + 2])
    from painlessinferenceacceleration_b200 import _lib as L
    rows = c._get_batch(qs, 64, 8, 0, 32, 'mix', [0] * len(qs), L.GET_HIER)
    for q, row in zip(qs[:64], rows[:64]):
        ids, m, sizes = c.hier_get(q, decoding_length=64, branch_length=8, min_output_size=32)
        assert ids == row[0] and np.array_equal(m, row[1]) and sizes == row[2]


def test_load_mem_of_a_reference_file_and_save_roundtrip(tmp_path):
    """load_mem reads a file written by the REFERENCE's save_mem (lookahead_cache.py:578-587) and answers like the
    reference did; save_mem -> load_mem into a fresh cache is lossless"""
    import json
    import os
    LookaheadCache = _gpu_cache_cls()
    c = LookaheadCache(eos_ids=[2], node_capacity=1 << 20)
    c.load_mem(os.path.join(R.GOLDEN, 'trie_mem_ref.json'))
    gets = json.load(open(os.path.join(R.GOLDEN, 'trie_mem_ref_gets.json')))
    for g in gets:
        ids, m, sizes = c.hier_get(g['q'], decoding_length=64, branch_length=8, min_output_size=32, mode='mix', idx=0)
        assert ids == g['ids'] and R.mask_rows(m) == g['mask'] and sizes == g['sizes'], g['q']
    # the loaded forest keeps learning
    c.put([3, 4, 5, 6, 7, 8], branch_length=9, mode='output', idx=-1)
    path = str(tmp_path / 'mem.json')
    c.save_mem(path)
    d = LookaheadCache(eos_ids=[2], node_capacity=1 << 20)
    d.load_mem(path)
    for g in gets + [{'q': [3, 4]}]:
        a = c.hier_get(g['q'], decoding_length=64, branch_length=8, min_output_size=32)
        b = d.hier_get(g['q'], decoding_length=64, branch_length=8, min_output_size=32)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert c.stats()['n_trees'] == d.stats()['n_trees']


def test_tree_methods_against_the_reference_tree():
    """the product's Tree (one-tree device trie): put / get / squeeze / reset_input_freq / n_node / n_output_node against
    the op stream recorded from the live reference Tree (tests/golden/gen_tree_methods_golden.py)"""
    from painlessinferenceacceleration_b200.common.lookahead_cache import Tree
    from tests.replay import replay_tree_methods
    assert replay_tree_methods(lambda tok, mn, mo: Tree(tok, max_node=mn, max_output_node=mo)) > 40


def test_pruned_walk_and_cluster_walk_match_the_full_walk(monkeypatch):
    """The three frequency walks of hier_get (pruned by the parent >= child count bound, full in one CTA, full in a
    thread-block cluster) give identical drafts on a forest with hot subtrees (tens of thousands of nodes below a
    frequent bigram), input-mode counts, squeezed trees and every mode / min-size combination incl. the python
    negative-index case (min size 0 -> the pruned walk must fall back)."""
    import bench
    LookaheadCache = _gpu_cache_cls()
    rng = np.random.default_rng(77)
    docs = bench.phrase_bank_prompts(300, 32000, seed=3)

    def build(prune, cluster):
        monkeypatch.setenv('PIA_TRIE_PRUNE', str(prune))
        monkeypatch.setenv('PIA_TRIE_GET_CLUSTER', str(cluster))
        c = LookaheadCache(eos_ids=[2], max_output_node=4096)
        for i, d in enumerate(docs):
            c.put(d, branch_length=9, mode='output', idx=-1)
            if i % 3 == 0:
                c.put(d[:64], branch_length=9, mode='input', idx=0)
        return c

    caches = [build(1, 1), build(0, 1), build(0, 8)]
    queries = [[3, 3], [3], [4, 3], [3, 4]]
    for _ in range(60):
        d = docs[int(rng.integers(0, len(docs)))]
        j = int(rng.integers(0, len(d) - 2))
        queries.append(d[j:j + 2])
    variants = [('mix', 0, 32), ('mix', 8, 16), ('mix', 0, 0), ('output', 0, 32), ('output', 0, 0), ('input', 16, 0),
                ('input', 0, 0), ('mix', 0, 64)]
    n_hot = 0
    for q in queries:
        for mode, mi, mo in variants:
            outs = []
            for c in caches:
                try:
                    ids, m, sizes = c.hier_get(list(q), decoding_length=64, branch_length=8, min_input_size=mi,
                                               min_output_size=mo, mode=mode, idx=0)
                    outs.append(([int(x) for x in ids], R.mask_rows(m), [int(x) for x in sizes]))
                except IndexError:
                    outs.append('IndexError')
            assert outs[0] == outs[1] == outs[2], f'{q} {mode} {mi} {mo}\n{outs[0]}\n{outs[1]}\n{outs[2]}'
            n_hot += outs[0] != 'IndexError' and len(outs[0][0]) >= 32
    assert n_hot > 50
    visited = [c.stats()['nodes_visited'] for c in caches]
    assert visited[0] * 2 < visited[1], visited      # the pruned walk really skips most of the hot subtrees
    for c in caches:
        assert c.stats()['error_flags'] == 0


def test_compact_reclaims_storage_and_changes_nothing_observable():
    """pia_trie_compact after squeezes and request resets: every query answers as before, the per-tree counters, the
    update lists' effects (a later squeeze) and further puts behave identically to an uncompacted twin, and the pools
    shrink."""
    LookaheadCache = _gpu_cache_cls()
    rng = np.random.default_rng(21)

    def stream(c, seed, n_req):
        r = np.random.default_rng(seed)
        for req in range(n_req):
            prompt = np.clip(r.zipf(1.3, size=int(r.integers(20, 120))), 3, 4999).tolist()
            c.put(prompt[1:], branch_length=9, final=False, mode='input', idx=req % 2)
            for _ in range(int(r.integers(2, 12))):
                c.stream_put(np.clip(r.zipf(1.3, size=int(r.integers(1, 9))), 3, 4999).tolist(), branch_length=9,
                             final=False, mode='output', idx=req % 2)
            c.stream_put([], branch_length=9, final=True, mode='output', idx=req % 2)

    # pools large enough for the UNcompacted twin to survive the whole test (it leaks ~0.8 M nodes per 1200 requests)
    twins = [LookaheadCache(eos_ids=[2], max_node=64, max_output_node=24, n_input_slots=2, node_capacity=1 << 22)
             for _ in range(2)]
    for c in twins:
        stream(c, 1, 1200)          # > 1024 touched trees: squeezes happen (tiny per-tree limits)
    before = twins[0].stats()
    nb, na = twins[0].compact()
    after = twins[0].stats()
    assert nb == before['nodes_used'] and na == after['nodes_used'] and na < 0.8 * nb, (nb, na)
    assert after['edges_used'] <= before['edges_used'] and after['n_trees'] == before['n_trees']
    assert after['error_flags'] == 0
    qs = [rng.integers(3, 40, size=2).tolist() for _ in range(300)] + [[3, 3], [4, 3], [3]]

    def answers():
        out = []
        for q in qs:
            rows = []
            for c in twins:
                ids, m, sizes = c.hier_get(list(q), decoding_length=64, branch_length=8, min_output_size=32)
                rows.append(([int(x) for x in ids], R.mask_rows(m), [int(x) for x in sizes]))
            assert rows[0] == rows[1], (q, rows)
            out.append(rows[0])
        return out

    a1 = answers()
    assert sum(len(r[0]) > 1 for r in a1) > 50
    for t in (3, 4, 5, 17):
        assert twins[0].tree_counters(t) == twins[1].tree_counters(t)
    for c in twins:                 # life goes on identically: more requests (and squeezes) on both
        stream(c, 2, 1100)
    answers()
    s0, s1 = twins[0].stats(), twins[1].stats()
    assert s0['error_flags'] == 0 and s1['error_flags'] == 0
    assert s0['n_trees'] == s1['n_trees'] and s0['nodes_used'] < s1['nodes_used']
    assert twins[0].maybe_compact(threshold=2.0) is None and twins[1].maybe_compact(threshold=0.0) is not None
    answers()
