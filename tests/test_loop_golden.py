# -*- coding: utf-8 -*-
"""Pins oracle/loop.py to results produced by the REFERENCE'S OWN loop code.

tests/golden/loop_*.npz were recorded by running the unmodified functions of
/root/reference/lookahead/lookahead/common/pretrained_model.py (lookahead_generation :947-1268,
lookahead_prepare_inputs_for_generation :666-756, _lookahead_update_model_kwargs_for_generation :764-892,
_update_cache :894-945) with the live reference trie (tests/golden/ref_loop.py, gen_loop_golden.py).  Here the
restated loop consumes the recorded logits step by step (so no floating-point noise enters) and must reproduce:
the draft the reference fed to the model, the accepted tokens, kv_idx / `continuous`, dls, edls, the stop rules
(max_length, eos), stop words, left padding, the repetition-penalty arithmetic, decoding_mode 'hier' and 'one'."""
import numpy as np
import pytest
import torch

from oracle.loop import _accept, lookahead_generate
from oracle.trie import OracleLookaheadCache
from tests import loop_golden as G


class ReplayBackend(object):
    """feeds the recorded logits of one request to the loop and checks what the loop asks of the model"""

    def __init__(self, meta, z, req):
        self.meta, self.z, self.req, self.i, self.P = meta, z, req, 0, 0
        self.compactions = []

    def rows(self):
        return self.P

    def forward(self, ids_in, m01, pos):
        st = self.req['steps'][self.i]
        self.i += 1
        if st['mask'] is None:
            assert ids_in[0].tolist() == self.req['prompt']
        else:
            assert ids_in[0].tolist() == st['decoding_ids'], (self.i, ids_in[0].tolist(), st['decoding_ids'])
            n = len(st['decoding_ids'])
            assert np.array_equal(m01[0, 0, :, -n:].numpy(), G.mask01(st))
            assert self.P == st['context_len'] - 1
        self.P += ids_in.shape[1]
        return G.step_logits(self.meta, self.z, st)

    def compact(self, keep_idx):
        st = self.req['steps'][self.i - 1]
        keep = keep_idx.tolist()
        ctx = st['context_len']
        assert keep[:ctx] == list(range(ctx))
        self.compactions.append(keep[ctx:])
        self.P = len(keep)


@pytest.mark.parametrize('name', G.names())
def test_oracle_loop_reproduces_the_reference_loop(name):
    meta, z = G.load(name)
    gen = meta['gen']
    eos = gen.get('eos_token_id', 2)
    trie = OracleLookaheadCache(eos_ids=[eos])
    n_multi = 0
    for req in meta['requests']:
        be = ReplayBackend(meta, z, req)
        am = None if req['attention_mask'] is None else torch.tensor([req['attention_mask']])
        out = lookahead_generate(None, trie, torch.tensor([req['prompt']]), max_new_tokens=req['max_new_tokens'],
                                 eos_token_id=[eos], decoding_length=meta['decoding_length'],
                                 branch_length=meta['branch_length'], decoding_mode=gen.get('decoding_mode', 'hier'),
                                 repetition_penalty=gen.get('repetition_penalty', 1.0),
                                 stop_words=set(gen['stop_words']) if 'stop_words' in gen else None, attention_mask=am,
                                 backend=be, trace=True)
        assert out['sequences'][0].tolist() == req['sequences']
        assert out['dls'] == req['dls'] and out['edls'] == req['edls']
        assert be.i == len(req['steps'])
        ci = 0
        for st, tr in zip(req['steps'], out['steps']):
            assert tr['tokens'] == st['tokens']
            if st['kv'] is not None:  # the reference compacted: same cache rows kept (:869)
                assert be.compactions[ci] == st['kv']['kv_idx']
                ci += 1
            n_multi += len(st['tokens']) > 1
        assert ci == len(be.compactions)
    assert n_multi >= 3  # the fixture does exercise multi-token accepts


@pytest.mark.parametrize('name', G.names())
def test_accept_routine_step_by_step(name):
    """oracle.loop._accept alone on every recorded (draft ids, tree mask, logits, context): tokens and logit indices"""
    meta, z = G.load(name)
    pen = meta['gen'].get('repetition_penalty', 1.0)
    for req in meta['requests']:
        for st in req['steps']:
            if st['mask'] is None:
                continue
            ctx = torch.tensor([req['sequences'][:st['context_len']]])
            toks, li = _accept(st['decoding_ids'], G.mask01(st), G.step_logits(meta, z, st), ctx, pen)
            assert toks == st['tokens']
            if st['kv'] is not None:
                assert [i - 1 + st['context_len'] for i in li[1:]] == st['kv']['kv_idx']
                assert st['kv']['continuous'] == (li[-1] == len(toks) - 1)


class BatchReplayBackend(object):
    """request `b` of a recorded batched call: returns the recorded logits of its row at every step it is active"""

    def __init__(self, meta, z, call, b):
        self.meta, self.z, self.call, self.b, self.P, self.t = meta, z, call, b, 0, 0

    def rows(self):
        return self.P

    def forward(self, ids_in, m01, pos):
        while True:
            st = self.call['steps'][self.t]
            self.t += 1
            if st['prefill'] or self.b in st['before']['batch_indices']:
                break
        lg = G.batch_step_logits(self.meta, self.z, st)
        if st['prefill']:
            row = self.b
        else:
            row = st['before']['batch_indices'].index(self.b)
            assert ids_in[0].tolist() == st['before']['ids'][row], (self.b, self.t)
            assert self.P == st['before']['cursors'][row]
        self.P += ids_in.shape[1]
        return lg[row:row + 1]

    def compact(self, keep_idx):
        self.P = keep_idx.numel()


@pytest.mark.parametrize('name', G.batch_names())
def test_oracle_batch_loop_reproduces_the_reference_batch_loop(name):
    """oracle/loop_batch.py against records of the reference's own batched loop (pretrained_model_batch.py:664-1330 +
    bat_get): per-slot trie puts, drafts of decoding_length // active // active nodes padded to the widest, bounded
    accept walks, requests leaving the batch at different steps, the padded output"""
    from oracle.loop_batch import lookahead_generate_batch
    meta, z = G.load_batch(name)
    gen = meta['gen']
    eos = gen.get('eos_token_id', 2)
    trie = OracleLookaheadCache(eos_ids=[eos])
    multi = 0
    for call in meta['calls']:
        ids = torch.tensor(call['input_ids'])
        out = lookahead_generate_batch(None, trie, ids, max_new_tokens=call['max_new_tokens'], eos_token_id=[eos],
                                       decoding_length=meta['decoding_length'], branch_length=meta['branch_length'],
                                       repetition_penalty=gen.get('repetition_penalty', 1.0),
                                       pad_token_id=gen.get('pad_token_id', 2),
                                       backend_factory=lambda b: BatchReplayBackend(meta, z, call, b), trace=True)
        assert out['sequences'].tolist() == call['sequences']
        assert out['dls'] == call['dls'] and out['edls'] == call['edls']
        rec = [st for st in call['steps'] if not st['prefill']]
        assert len(rec) == len(out['steps'])
        for st, tr in zip(rec, out['steps']):
            assert tr['active'] == st['before']['batch_indices'] and tr['ids'] == st['before']['ids']
            assert tr['tokens'] == st['tokens']
        multi += sum(e > 1 for e in call['edls'])
    assert multi >= 3
