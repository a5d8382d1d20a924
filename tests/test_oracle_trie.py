# -*- coding: utf-8 -*-
"""Pins the CPU oracle (oracle/trie_oracle.c) to the reference:
 * the reference's own golden vectors (lookahead/tests/test_lookahead_cache.py:16-45),
 * SURVEY.md appendix A.1, and
 * op streams recorded from the live reference module (tests/golden/gen_trie_golden.py)."""
import numpy as np
import pytest

from oracle.trie import OracleLookaheadCache, OracleTree
from tests import replay as R

STREAMS = ['trie_survey_a1.json', 'trie_small_v12.json', 'trie_small_v6_stop.json', 'trie_small_v30_batch.json',
           'trie_small_v8_dl128.json', 'trie_squeeze.json', 'trie_zipf.json']


def _stats(c):
    return {'n_trees': c.n_trees(), 'total_nodes': c.total_nodes(), 'n_update_trees': c.n_update_trees()}


def test_reference_unit_vectors_literal():
    # bodies of lookahead/tests/test_lookahead_cache.py:16-45 with the oracle's Tree
    tree = OracleTree(1)
    tree.put([1, 2, 3, 4], mode='output', idx=-1)
    ids, mask, sizes = tree.get([1], max_size=63, max_length=8, min_input_size=0, min_output_size=0,
                                output_weight=1e-4, mode='mix', idx=0)
    assert ids == [1, 2, 3, 4]
    assert mask.shape == (4, 4)
    assert np.sum(np.abs(mask - np.array([[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1]]))) == 0
    tree = OracleTree(1)
    tree.put([1, 2, 3], mode='output', idx=-1)
    tree.put([1, 2, 4], mode='output', idx=-1)
    ids, mask, sizes = tree.get([1], max_size=63, max_length=8, min_input_size=0, min_output_size=0,
                                output_weight=1e-4, mode='mix', idx=0)
    assert ids == [1, 2, 3, 4]
    assert np.sum(np.abs(mask - np.array([[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 0, 1]]))) == 0


def test_reference_unit_vectors_fixture():
    for case in R.load('trie_unit.json'):
        t = OracleTree(case['tree_token'])
        for p in case['puts']:
            t.put(p, mode='output', idx=-1)
        g = case['get']
        ids, m, sizes = t.get(g['q'], max_size=g['max_size'], max_length=g['max_length'],
                              min_input_size=g['min_input_size'], min_output_size=g['min_output_size'],
                              mode=g['mode'], idx=g['idx'])
        assert {'ids': ids, 'mask': R.mask_rows(m), 'sizes': sizes} == case['out']


@pytest.mark.parametrize('name', STREAMS)
def test_recorded_streams(name):
    fx = R.load(name)
    cache = R.make_cache(OracleLookaheadCache, fx['ctor'])
    n = R.replay(cache, fx['ops'], stats_fn=_stats, tag=name)
    assert n > 0


def test_tree_methods_against_the_reference_tree():
    """Tree.put / get / squeeze / reset_input_freq / counters (lookahead_cache.py:24-333) recorded from the live reference"""
    from oracle.trie import OracleTree
    from tests.replay import replay_tree_methods
    assert replay_tree_methods(lambda tok, mn, mo: OracleTree(tok, max_node=mn, max_output_node=mo)) > 40
