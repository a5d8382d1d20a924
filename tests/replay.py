# -*- coding: utf-8 -*-
"""Replays a recorded op stream (tests/golden/trie_*.json, produced by the live reference trie) against any
object with the LookaheadCache method surface and checks every recorded result bit-for-bit."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def mask_rows(m):
    m = np.asarray(m)
    return [int(sum(int(v) << j for j, v in enumerate(row.astype(np.int64).tolist()))) for row in m]


def make_cache(cls, ctor):
    kw = dict(ctor)
    if 'stop_words' in kw:
        kw['stop_words'] = set(kw['stop_words'])
    return cls(**kw)


def replay(cache, ops, stats_fn=None, tag=''):
    """returns the number of checked get results"""
    checked = 0
    for k, op in enumerate(ops):
        name = op[0]
        where = f'{tag} op#{k} {op[:8]}'
        if name == 'put':
            _, ids, bl, final, mode, idx = op
            cache.put(list(ids), branch_length=bl, final=final, mode=mode, idx=idx)
        elif name == 'stream_put':
            _, ids, bl, final, mode, idx = op
            cache.stream_put(list(ids), branch_length=bl, final=final, mode=mode, idx=idx)
        elif name in ('hier_get', 'one_get', 'par_get'):
            _, q, dl, bl, mi, mo, mode, idx, want = op
            try:
                ids, m, sizes = getattr(cache, name)(list(q), decoding_length=dl, branch_length=bl, min_input_size=mi,
                                                     min_output_size=mo, mode=mode, idx=idx)
                got = {'ids': [int(x) for x in ids], 'mask': mask_rows(m), 'sizes': [int(x) for x in sizes]}
            except IndexError:
                got = {'err': 'IndexError'}
            assert got == want, f'{where}\n got={got}\nwant={want}'
            checked += 1
        elif name == 'bat_get':
            _, qs, dl, bl, cursors, mode, indices, dmode, want = op
            ids, m, sizes = cache.bat_get([list(q) for q in qs], decoding_length=dl, branch_length=bl,
                                          decoding_cursors=list(cursors), mode=mode, indices=list(indices),
                                          decoding_mode=dmode)
            got = {'ids': [[int(v) for v in x] for x in ids], 'mask': np.asarray(m).astype(int).tolist(),
                   'sizes': [[int(v) for v in s] for s in sizes]}
            assert got == want, f'{where}\n got={got}\nwant={want}'
            checked += 1
        elif name == 'fresh':
            cache.fresh()
        elif name == 'reset_input_freqs':
            cache.reset_input_freqs(op[1])
        elif name == 'squeeze_branch_counts':
            cache.squeeze_branch_counts()
        elif name == 'set_stop_words':
            cache.stop_words = set(op[1])
        elif name == 'set_eos':
            cache.eos_ids = list(op[1])
        elif name == 'stats':
            if stats_fn is not None:
                got = stats_fn(cache)
                for key, val in op[1].items():
                    if key in got:
                        assert got[key] == val, f'{where} stats {key}: got {got[key]} want {val}'
        else:
            raise ValueError(name)
    return checked


def replay_tree_methods(make_tree):
    """tests/golden/trie_tree_methods.json (recorded from the live reference Tree) against `make_tree(token, max_node,
    max_output_node)`; returns the number of checked results"""
    checked = 0
    for case in load('trie_tree_methods.json'):
        t = make_tree(case['token'], case['max_node'], case['max_output_node'])
        for k, op in enumerate(case['ops']):
            if op[0] == 'put':
                t.put(list(op[1]), mode=op[2], idx=op[3])
            elif op[0] == 'get':
                _, q, mode, idx, want = op
                try:
                    ids, m, sizes = t.get(list(q), max_size=16, max_length=6, min_input_size=0, min_output_size=4, mode=mode,
                                          idx=idx)
                    got = {'ids': [int(x) for x in ids], 'mask': mask_rows(m), 'sizes': [int(x) for x in sizes]}
                except IndexError:
                    got = {'err': 'IndexError'}
                assert got == want, f'op#{k} {op[:4]}\n got={got}\nwant={want}'
                checked += 1
            elif op[0] == 'squeeze':
                t.squeeze()
                assert (int(t.n_node), int(t.n_output_node)) == (op[1], op[2]), f'op#{k} squeeze counters'
                checked += 1
            elif op[0] == 'reset_input_freq':
                t.reset_input_freq(op[1])
            elif op[0] == 'counters':
                assert (int(t.n_node), int(t.n_output_node)) == (op[1], op[2]), f'op#{k} counters'
    return checked
