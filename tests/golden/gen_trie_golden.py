# -*- coding: utf-8 -*-
"""Generates tests/golden/trie_*.json by driving the LIVE reference trie
(/root/reference/lookahead/lookahead/common/lookahead_cache.py, numpy only) with seeded op streams.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONPATH=/root/reference/lookahead python tests/golden/gen_trie_golden.py

Every op is recorded together with what the reference returned, so the fixtures pin both the CPU oracle
(tests/test_oracle_trie.py) and the CUDA trie (tests/test_gpu_trie.py) without the reference at run time.
Mask rows are stored as integers (bit j of row i == mask[i, j]).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference/lookahead')
from lookahead.common.lookahead_cache import LookaheadCache, Tree  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mask_rows(m):
    m = np.asarray(m)
    return [int(sum(int(v) << j for j, v in enumerate(row.astype(np.int64).tolist()))) for row in m]


def count_nodes(nodes):
    s = len(nodes)
    for n in nodes.values():
        s += count_nodes(n.children)
    return s


class Recorder(object):
    def __init__(self, **ctor):
        self.ctor = dict(ctor)
        kw = dict(ctor)
        if 'stop_words' in kw:
            kw['stop_words'] = set(kw['stop_words'])
        self.c = LookaheadCache(**kw)
        self.ops = []

    def put(self, ids, bl=8, final=False, mode='output', idx=0):
        self.c.put(list(ids), branch_length=bl, final=final, mode=mode, idx=idx)
        self.ops.append(['put', list(map(int, ids)), bl, bool(final), mode, idx])

    def stream_put(self, ids, bl=8, final=False, idx=0):
        self.c.stream_put(list(ids), branch_length=bl, final=final, mode='output', idx=idx)
        self.ops.append(['stream_put', list(map(int, ids)), bl, bool(final), 'output', idx])

    def _get(self, name, q, dl, bl, mi, mo, mode, idx):
        try:
            ids, m, sizes = getattr(self.c, name)(list(q), decoding_length=dl, branch_length=bl, min_input_size=mi,
                                                  min_output_size=mo, mode=mode, idx=idx)
            out = {'ids': list(map(int, ids)), 'mask': mask_rows(m), 'sizes': list(map(int, sizes))}
        except IndexError:
            out = {'err': 'IndexError'}
        self.ops.append([name, list(map(int, q)), dl, bl, mi, mo, mode, idx, out])
        return out

    def hier_get(self, q, dl=64, bl=8, mi=0, mo=0, mode='mix', idx=0):
        return self._get('hier_get', q, dl, bl, mi, mo, mode, idx)

    def one_get(self, q, dl=64, bl=8, mi=0, mo=0, mode='mix', idx=0):
        return self._get('one_get', q, dl, bl, mi, mo, mode, idx)

    def par_get(self, q, dl=16, bl=8, mi=0, mo=0, mode='mix', idx=0):
        return self._get('par_get', q, dl, bl, mi, mo, mode, idx)

    def bat_get(self, qs, dl, bl, cursors, mode, indices, decoding_mode='hier'):
        ids, m, sizes = self.c.bat_get([list(q) for q in qs], decoding_length=dl, branch_length=bl,
                                       decoding_cursors=list(cursors), mode=mode, indices=list(indices),
                                       decoding_mode=decoding_mode)
        out = {'ids': [list(map(int, x)) for x in ids], 'mask': np.asarray(m).astype(int).tolist(),
               'sizes': [list(map(int, s)) for s in sizes]}
        self.ops.append(['bat_get', [list(map(int, q)) for q in qs], dl, bl, list(cursors), mode, list(indices),
                         decoding_mode, out])

    def simple(self, name, *args):
        getattr(self.c, name)(*args)
        self.ops.append([name] + list(args))

    def set_stop_words(self, words):
        self.c.stop_words = set(words)
        self.ops.append(['set_stop_words', sorted(words)])

    def set_eos(self, eos):
        self.c.eos_ids = list(eos)
        self.ops.append(['set_eos', list(eos)])

    def stats(self):
        total = sum(count_nodes(t.nodes) for t in self.c.mem.values())
        self.ops.append(['stats', {'n_trees': len(self.c.mem), 'total_nodes': int(total),
                                   'n_update_trees': len(self.c._update_trees)}])

    def dump(self, name):
        path = os.path.join(HERE, name)
        with open(path, 'w') as f:
            json.dump({'ctor': self.ctor, 'ops': self.ops}, f, separators=(',', ':'))
        print(name, len(self.ops), 'ops', os.path.getsize(path) // 1024, 'KiB')


def gen_unit():
    """the reference's own golden vectors (lookahead/tests/test_lookahead_cache.py:16-45) at Tree level,
    plus SURVEY.md appendix A.1."""
    cases = []
    for puts in ([[1, 2, 3, 4]], [[1, 2, 3], [1, 2, 4]]):
        t = Tree(1)
        for p in puts:
            t.put(p, mode='output', idx=-1)
        ids, m, sizes = t.get([1], max_size=63, max_length=8, min_input_size=0, min_output_size=0,
                              output_weight=1e-4, mode='mix', idx=0)
        cases.append({'tree_token': 1, 'puts': puts, 'get': {'q': [1], 'max_size': 63, 'max_length': 8,
                                                              'min_input_size': 0, 'min_output_size': 0,
                                                              'mode': 'mix', 'idx': 0},
                      'out': {'ids': list(map(int, ids)), 'mask': mask_rows(m), 'sizes': list(map(int, sizes))}})
    with open(os.path.join(HERE, 'trie_unit.json'), 'w') as f:
        json.dump(cases, f)
    r = Recorder(eos_ids=[2])
    for p in ([10, 11, 20, 21, 22], [10, 11, 20, 21, 22], [10, 11, 20, 23], [10, 11, 30, 31]):
        r.put(p, bl=9, mode='output', idx=-1)
    out = r.hier_get([10, 11], dl=64, bl=8, mo=32)
    assert out['ids'] == [11, 20, 21, 22, 23, 30, 31] and out['mask'] == [0x01, 0x03, 0x07, 0x0f, 0x13, 0x21, 0x61]
    r.one_get([10, 11])
    r.par_get([10, 11], dl=16)
    r.hier_get([99, 98])
    r.hier_get([10, 11], dl=1)
    r.hier_get([10, 11], bl=0)
    r.hier_get([11])
    r.hier_get([10])
    r.hier_get([10, 0])
    r.stats()
    r.dump('trie_survey_a1.json')


def gen_small_vocab(seed, name, V, n_req, stop_words=(), eos=(2,), max_output_node=512, with_batch=False,
                    dl=64, bl=8):
    """request-shaped stream on a tiny vocabulary: over-full subtrees, every mode, prompt reset, stream_put."""
    rng = np.random.default_rng(seed)
    r = Recorder(eos_ids=list(eos), stop_words=sorted(stop_words), max_output_node=max_output_node)
    for req in range(n_req):
        plen = int(rng.integers(2, 40))
        prompt = rng.integers(0, V, size=plen).tolist()
        idx = int(rng.integers(0, 3)) if with_batch else 0
        r.put(prompt[1:], bl=bl + 1, mode='input', idx=idx)
        seq = list(prompt)
        steps = int(rng.integers(1, 14))
        for s in range(steps):
            q = seq[-2:]
            mo = int(rng.choice([0, 1, dl // 2, dl // 2, dl // 2]))
            mode = str(rng.choice(['mix', 'mix', 'mix', 'input', 'output']))
            mi = int(rng.choice([0, 0, 0, 3]))
            r.hier_get(q, dl=dl, bl=int(rng.choice([bl, bl, 3, 1])), mi=mi, mo=mo, mode=mode, idx=idx)
            if rng.random() < 0.15:
                r.one_get(q, dl=dl, bl=bl, mode=mode, idx=idx)
            if rng.random() < 0.15:
                r.par_get(q, dl=16, bl=bl, mo=8, mode='mix', idx=idx)
            if rng.random() < 0.1:
                r.hier_get(seq[-3:], dl=dl, bl=bl, mo=dl // 2, idx=idx)
            k = int(rng.integers(1, bl + 2))
            new = rng.integers(0, V, size=k).tolist()
            seq.extend(new)
            r.stream_put(new, bl=bl + 1, final=False, idx=idx)
            if any(e in new for e in eos):
                break
        r.stream_put([], bl=bl + 1, final=True, idx=idx)
        if req % 7 == 3:
            r.put(rng.integers(0, V, size=int(rng.integers(2, 30))).tolist(), bl=bl + 1, mode='output', idx=-1)
        if with_batch and req % 5 == 4:
            qs = [rng.integers(0, V, size=2).tolist() for _ in range(3)]
            r.bat_get(qs, dl, bl, [int(x) for x in rng.integers(5, 12, size=3)], 'mix', [0, 1, 2], 'hier')
        if req % 11 == 10:
            r.stats()
    r.stats()
    r.dump(name)


def gen_squeeze(seed, name):
    """>=1024 distinct trees touched so that squeeze_branch_counts fires; tiny max_output_node so trees squeeze."""
    rng = np.random.default_rng(seed)
    r = Recorder(eos_ids=[2], max_output_node=6, max_node=40)
    V = 1400
    for rep in range(3):
        for base in range(3, V, 97):
            doc = [(base + 7 * j) % (V - 3) + 3 for j in range(97)]
            r.put(doc, bl=5, mode='output', idx=-1)
            if rep == 2 and base % 2:
                r.put(doc[:20], bl=5, mode='output', idx=-1)
        hot = rng.integers(3, 40, size=60).tolist()
        r.put(hot, bl=5, mode='output', idx=-1)
        r.put(hot[:30], bl=5, mode='input', idx=0)
        r.stats()
        for q in ([hot[0], hot[1]], [hot[5], hot[6]], [3, 10], [10, 17]):
            r.hier_get(q, dl=16, bl=4, mo=8)
        r.stream_put(hot[:12], bl=5, final=False, idx=0)
        r.stream_put([], bl=5, final=True, idx=0)   # reset + squeeze (>=1024 trees touched)
        r.stats()
        for q in ([hot[0], hot[1]], [hot[5], hot[6]], [3, 10], [10, 17], [hot[2]]):
            r.hier_get(q, dl=16, bl=4, mo=8)
            r.hier_get(q, dl=64, bl=8, mo=32)
    r.simple('fresh')
    r.stats()
    r.hier_get([3, 10])
    r.put([3, 10, 17, 24, 31], bl=5, mode='output', idx=-1)
    r.hier_get([3, 10])
    r.stats()
    r.dump(name)


def gen_zipf(seed, name, n_docs=30, n_req=12):
    """the benchmark's phrase-bank stream at fixture scale, V=32000."""
    rng = np.random.default_rng(seed)
    V = 32000
    bank = [np.clip(rng.zipf(1.3, size=int(rng.integers(4, 25))), 3, V - 1).tolist() for _ in range(200)]

    def doc(n):
        out = []
        while len(out) < n:
            out.extend(bank[int(rng.integers(0, len(bank)))])
        return out[:n]

    r = Recorder(eos_ids=[2])
    for _ in range(n_docs):
        r.put(doc(128), bl=9, mode='output', idx=-1)
    r.stats()
    for _ in range(n_req):
        prompt = doc(64)
        r.put(prompt[1:], bl=9, mode='input', idx=0)
        seq = list(prompt)
        cont = doc(48)
        pos = 0
        while pos < len(cont):
            r.hier_get(seq[-2:], dl=64, bl=8, mo=32)
            k = int(rng.integers(1, 5))
            new = cont[pos:pos + k]
            pos += k
            seq.extend(new)
            r.stream_put(new, bl=9, final=False, idx=0)
        r.stream_put([], bl=9, final=True, idx=0)
    r.stats()
    r.dump(name)


def gen_mem_file():
    """a trie persisted by the REFERENCE's save_mem (:578-582) plus what it answers, for load_mem parity"""
    rng = np.random.default_rng(9)
    c = LookaheadCache(eos_ids=[2])
    for _ in range(40):
        c.put(rng.integers(3, 14, size=int(rng.integers(4, 40))).tolist(), branch_length=9, mode='output', idx=-1)
    c.put(rng.integers(3, 14, size=30).tolist(), branch_length=9, mode='input', idx=0)
    c.save_mem(os.path.join(HERE, 'trie_mem_ref.json'))
    qs = [rng.integers(3, 14, size=2).tolist() for _ in range(25)]
    outs = []
    for q in qs:
        ids, m, sizes = c.hier_get(q, decoding_length=64, branch_length=8, min_output_size=32, mode='mix', idx=0)
        outs.append({'q': q, 'ids': list(map(int, ids)), 'mask': mask_rows(m), 'sizes': list(map(int, sizes))})
    with open(os.path.join(HERE, 'trie_mem_ref_gets.json'), 'w') as f:
        json.dump(outs, f)
    print('trie_mem_ref.json', os.path.getsize(os.path.join(HERE, 'trie_mem_ref.json')) // 1024, 'KiB')


if __name__ == '__main__':
    gen_unit()
    gen_mem_file()
    gen_small_vocab(1, 'trie_small_v12.json', V=12, n_req=40)
    gen_small_vocab(2, 'trie_small_v6_stop.json', V=6, n_req=30, stop_words=(3, 4), eos=(2, 5))
    gen_small_vocab(3, 'trie_small_v30_batch.json', V=30, n_req=30, with_batch=True, dl=48, bl=6)
    gen_small_vocab(4, 'trie_small_v8_dl128.json', V=8, n_req=16, dl=128, bl=12)
    gen_squeeze(5, 'trie_squeeze.json')
    gen_zipf(6, 'trie_zipf.json')
