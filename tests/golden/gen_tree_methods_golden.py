# -*- coding: utf-8 -*-
"""tests/golden/trie_tree_methods.json: the per-tree maintenance methods of the LIVE reference Tree
(lookahead_cache.py:24-333) - put / get / squeeze (:295-318) / reset_input_freq (:320-333) / n_node / n_output_node -
on seeded op streams.  Build container only:  python tests/golden/gen_tree_methods_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference/lookahead')
from lookahead.common.lookahead_cache import Tree  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mask_rows(m):
    m = np.asarray(m)
    return [int(sum(int(v) << j for j, v in enumerate(row.astype(np.int64).tolist()))) for row in m]


def main():
    rng = np.random.default_rng(11)
    cases = []
    for case in range(4):
        max_node, max_out = (40, 12) if case % 2 == 0 else (25, 1000)
        t = Tree(5, max_node=max_node, max_output_node=max_out)
        ops = []
        for _ in range(60):
            r = rng.random()
            if r < 0.55:
                ids = rng.integers(3, 9, size=int(rng.integers(1, 7))).tolist()
                mode = 'output' if rng.random() < 0.6 else 'input'
                idx = -1 if mode == 'output' else int(rng.integers(0, 2))
                t.put(list(ids), mode=mode, idx=idx)
                ops.append(['put', ids, mode, idx])
            elif r < 0.8:
                q = rng.integers(3, 9, size=int(rng.integers(0, 3))).tolist()
                mode = ['mix', 'output', 'input'][int(rng.integers(0, 3))]
                idx = int(rng.integers(0, 2))
                try:
                    ids, m, sizes = t.get(list(q), max_size=16, max_length=6, min_input_size=0, min_output_size=4,
                                          mode=mode, idx=idx)
                    out = {'ids': [int(x) for x in ids], 'mask': mask_rows(m), 'sizes': [int(x) for x in sizes]}
                except IndexError:
                    out = {'err': 'IndexError'}
                ops.append(['get', q, mode, idx, out])
            elif r < 0.9:
                t.squeeze()
                ops.append(['squeeze', int(t.n_node), int(t.n_output_node)])
            else:
                idx = int(rng.integers(0, 2))
                t.reset_input_freq(idx)
                ops.append(['reset_input_freq', idx])
            ops.append(['counters', int(t.n_node), int(t.n_output_node)])
        cases.append(dict(token=5, max_node=max_node, max_output_node=max_out, ops=ops))
    with open(os.path.join(HERE, 'trie_tree_methods.json'), 'w') as f:
        json.dump(cases, f)
    print(sum(len(c['ops']) for c in cases), 'ops;', sum(1 for c in cases for o in c['ops'] if o[0] == 'squeeze'), 'squeezes')


if __name__ == '__main__':
    main()
