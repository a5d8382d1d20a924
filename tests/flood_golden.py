# -*- coding: utf-8 -*-
"""loader / replayer for tests/golden/flood_draft.npz (recorded from the reference's Triton kernels run under the
Triton interpreter, tests/golden/gen_flood_golden.py)"""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flood_draft.npz')


def load():
    z = np.load(PATH)
    return json.loads(bytes(z['ops']).decode()), z


def replay(impl, tags=('sq', 'wide')):
    """drives `impl` (an object with update / retrieve / verify / cache_move / tables) through the recorded stream and
    yields (description, got, want) triples"""
    ops, z = load()
    for tag in tags:
        T, BL, BC, V = (int(x) for x in z[f'{tag}_geom'])
        impl.reset(T, BL, BC, V)
        for op in ops:
            if op[0] != tag:
                continue
            k = op[2]
            if op[1] == 'update':
                impl.update(z[f'{tag}_{k}_tokens'].tolist())
            elif op[1] == 'retrieve':
                out, masks = impl.retrieve(z[f'{tag}_{k}_queries'].tolist(), op[3])
                yield (tag, k, 'retrieve tokens'), np.asarray(out), z[f'{tag}_{k}_out']
                yield (tag, k, 'retrieve masks'), np.asarray(masks), z[f'{tag}_{k}_masks']
            else:
                RC, bs = op[3], op[4]
                o, s, d = impl.verify(z[f'{tag}_{k}_input'], z[f'{tag}_{k}_next'], z[f'{tag}_{k}_offs'], bs, RC)
                yield (tag, k, 'verify out'), np.asarray(o), z[f'{tag}_{k}_vout']
                yield (tag, k, 'verify src'), np.asarray(s), z[f'{tag}_{k}_vsrc']
                yield (tag, k, 'verify dst'), np.asarray(d), z[f'{tag}_{k}_vdst']
                cache = np.arange(400 * 6, dtype=np.float32).reshape(400, 6).copy()
                yield (tag, k, 'cache'), np.asarray(impl.cache_move(cache, z[f'{tag}_{k}_vsrc'], z[f'{tag}_{k}_vdst'])), \
                    z[f'{tag}_{k}_cache']
        freq, table = impl.tables()
        yield (tag, 'final', 'freq'), np.asarray(freq), z[f'{tag}_final_freq']
        yield (tag, 'final', 'table'), np.asarray(table), z[f'{tag}_final_table']
