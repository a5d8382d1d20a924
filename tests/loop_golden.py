# -*- coding: utf-8 -*-
"""loader for tests/golden/loop_*.npz (written by tests/golden/gen_loop_golden.py from the reference's own loop code)"""
import glob
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def names():
    return sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLD, 'loop_*.npz')))


def load(name):
    z = np.load(os.path.join(GOLD, f'loop_{name}.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    return meta, z


def step_logits(meta, z, step):
    """recorded logits of one verify step as a torch tensor [1, rows, V] in the recorded dtype"""
    import torch
    a = z[step['logits']]
    if meta['dtype'] == 'bfloat16':
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)[None]
    return torch.from_numpy(a.copy())[None]


def step_mask(step):
    """uint64 ancestor rows of the step's draft (None for the prefill step)"""
    if step['mask'] is None:
        return None
    return np.array([int(v) for v in step['mask']], dtype=np.uint64)


def mask01(step):
    rows = step_mask(step)
    n = len(rows)
    return np.array([[(int(r) >> j) & 1 for j in range(n)] for r in rows], dtype=np.int64)


def batch_names():
    return sorted(os.path.basename(f)[10:-4] for f in glob.glob(os.path.join(GOLD, 'batchloop_*.npz')))


def load_batch(name):
    z = np.load(os.path.join(GOLD, f'batchloop_{name}.npz'))
    return json.loads(bytes(z['meta']).decode()), z


def batch_step_logits(meta, z, step):
    """[k, n, V] tensor of one recorded batched verify step (k = active requests in batch_indices order)"""
    import torch
    a = z[step['logits']]
    if meta['dtype'] == 'bfloat16':
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    return torch.from_numpy(a.copy())
