# -*- coding: utf-8 -*-
"""The CUDA accept / KV-compaction kernels and the fused DEVICE LOOP against results of the reference's own loop code.

tests/golden/loop_*.npz (tests/golden/gen_loop_golden.py) hold, for every verify step of requests run through the
unmodified reference functions (pretrained_model.py:666-1268) with the live reference trie: the context, the draft the
reference fed to the model, the logits it got back and what it accepted.  Two checks, both through the C ABI:
  * step by step: pia_accept + pia_kv_compact on the recorded (ids, mask, logits, context) give the recorded tokens,
    logit indices (-> kv_idx, :869) and cache rows, and raise `finished` exactly on the recorded last step;
  * the whole device loop (GPU trie get -> [recorded logits] -> accept -> compaction -> stream_put, one CUDA graph per
    step, tries carried across requests) with the verify forward replaced by the recorded logits reproduces the
    reference's drafts, tokens, dls and edls for every request - bit exact, no floating point in between."""
import types

import numpy as np
import pytest
import torch
from torch import nn

from tests import loop_golden as G

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BF16 = [n for n in G.names() if 'bf16' in n]


def _pad_logits(meta, z, st, rows=64):
    lg = G.step_logits(meta, z, st)[0]
    out = torch.zeros((rows, lg.shape[1]), dtype=torch.bfloat16)
    out[:lg.shape[0]] = lg
    return out


@pytest.mark.parametrize('name', BF16)
def test_accept_and_compaction_kernels_step_by_step(name):
    from painlessinferenceacceleration_b200.common import ops
    meta, z = G.load(name)
    gen = meta['gen']
    eos, pen, V = gen.get('eos_token_id', 2), gen.get('repetition_penalty', 1.0), meta['vocab']
    dev = torch.device(DEV)
    i32 = dict(dtype=torch.int32, device=dev)
    n_checked = n_moved = 0
    for req in meta['requests']:
        max_length = len(req['prompt']) + req['max_new_tokens']
        max_seq = max_length + 80
        acc = ops.Accept(V, 64, pen, [eos], max_length, dev)
        pad_len = 0 if req['attention_mask'] is None else req['attention_mask'].index(1)
        for si, st in enumerate(req['steps']):
            if st['mask'] is None:
                continue
            n, ctx = len(st['decoding_ids']), st['context_len']
            logits = _pad_logits(meta, z, st).to(dev)
            ids = torch.zeros((64,), **i32)
            ids[:n] = torch.tensor(st['decoding_ids'], **i32)
            mask = torch.zeros((64, 1), dtype=torch.int64, device=dev)
            mask[:n, 0] = torch.from_numpy(G.step_mask(st).view(np.int64)).to(dev)
            seq = torch.zeros((max_seq + 8,), **i32)
            seq[:ctx] = torch.tensor(req['sequences'][:ctx], **i32)
            seq_len, prefix = torch.tensor([ctx], **i32), torch.tensor([ctx - 1], **i32)
            dn, fin = torch.tensor([n], **i32), torch.zeros((1,), **i32)
            toks, cnt, nodes = torch.zeros((64,), **i32), torch.zeros((1,), **i32), torch.zeros((64,), **i32)
            # marker cache: row r holds (r // 128, r % 128) - exact in bf16
            r = torch.arange(max_seq, device=dev)
            kc = torch.zeros((1, 1, max_seq, 128), dtype=torch.bfloat16, device=dev)
            kc[0, 0, :, 0], kc[0, 0, :, 1] = (r // 128).to(torch.bfloat16), (r % 128).to(torch.bfloat16)
            vc = kc.clone()
            before = kc.clone()
            acc.run(logits, ids, mask, dn, seq, seq_len, pad_len, toks, cnt, nodes, prefix, fin)
            ops.kv_compact(kc, vc, nodes, cnt, prefix)
            torch.cuda.synchronize()
            c = int(cnt)
            assert toks[:c].tolist() == st['tokens'], (name, si)
            assert int(seq_len) == ctx + c and int(prefix) == ctx - 1 + c
            assert seq[ctx:ctx + c].tolist() == st['tokens']
            li = nodes[:c].tolist()
            assert li[0] == 0
            kv_idx = [j - 1 + ctx for j in li[1:]]           # pretrained_model.py:869
            if st['kv'] is not None:
                assert kv_idx == st['kv']['kv_idx'] and st['kv']['continuous'] == (li[-1] == c - 1)
                n_moved += not st['kv']['continuous']
            else:                                            # every draft node accepted: nothing to move (:865)
                assert c - 1 == n - 1 and li == list(range(c))
            # cache rows after compaction = concat(rows[:ctx], rows[kv_idx]) (:904-905)
            keep = list(range(ctx)) + kv_idx
            assert torch.equal(kc[0, 0, :len(keep)], before[0, 0, keep]) and torch.equal(vc[0, 0, :len(keep)], before[0, 0, keep])
            last = si == len(req['steps']) - 1
            assert bool(int(fin)) == last, (name, si)
            n_checked += 1
    assert n_checked > 20 and n_moved >= 1


def _replay_model(vocab, dev):
    from painlessinferenceacceleration_b200.common.pretrained_model import LookaheadPreTrainedModel

    class Replay(LookaheadPreTrainedModel):
        """the product's loop driver with the verify forward replaced by recorded logits"""

        def __init__(self):
            super().__init__(types.SimpleNamespace(eos_token_id=2, pad_token_id=0))
            self.anchor = nn.Parameter(torch.zeros(1, device=dev), requires_grad=False)
            self.cap = 96
            self.all = torch.zeros((self.cap, 64, vocab), dtype=torch.bfloat16, device=dev)
            self.ids_log = torch.zeros((self.cap, 64), dtype=torch.int32, device=dev)
            self.n_log = torch.zeros((self.cap,), dtype=torch.int32, device=dev)
            self.mask_log = torch.zeros((self.cap, 64), dtype=torch.int64, device=dev)
            self.step = torch.zeros((1,), dtype=torch.int64, device=dev)
            self.first = torch.zeros((1, vocab), dtype=torch.bfloat16, device=dev)

        def geometry(self):
            return dict(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, head_dim=128, inter=128, vocab=vocab)

        def rope_tables(self, max_pos):
            z = torch.zeros((max_pos, 64), dtype=torch.bfloat16, device=dev)
            return z, z.clone()

        def load(self, meta, z, req):
            steps = req['steps']
            assert len(steps) - 1 <= self.cap
            self.all.zero_()
            for k, st in enumerate(steps[1:]):
                self.all[k] = _pad_logits(meta, z, st).to(dev)
            self.first[0] = G.step_logits(meta, z, steps[0])[0, -1].to(dev)
            self.step.zero_()
            self.ids_log.zero_()

        def _prefill_kv(self, rt, prompt_len):
            if not hasattr(rt, 'chain'):
                rt.chain = rt.chain_mask_rows()
            rt.logits[0:1] = self.first
            rt.prefix_len.fill_(prompt_len)

        def _verify_layers(self, rt, bufs=None, last_only=False):
            rt.logits.copy_(self.all.index_select(0, self.step)[0])
            self.ids_log.index_copy_(0, self.step, rt.ids)
            self.n_log.index_copy_(0, self.step, rt.n)
            self.mask_log.index_copy_(0, self.step, rt.mask[:, :, 0])
            self.step += 1

    return Replay()


@pytest.mark.parametrize('name', BF16)
def test_device_loop_reproduces_the_reference_loop(name):
    from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache
    meta, z = G.load(name)
    gen = meta['gen']
    eos, V = gen.get('eos_token_id', 2), meta['vocab']
    dev = torch.device(DEV)
    model = _replay_model(V, dev)
    model.lookahead_cache = LookaheadCache(eos_ids=[eos], device=dev, vocab_capacity=1024, node_capacity=1 << 20)
    multi = 0
    for ri, req in enumerate(meta['requests']):
        model.load(meta, z, req)
        dk = {'use_lookahead': True, 'decoding_length': meta['decoding_length'], 'branch_length': meta['branch_length'],
              'decoding_mode': gen.get('decoding_mode', 'hier')}
        if 'stop_words' in gen:
            dk['stop_words'] = set(gen['stop_words'])
        am = None if req['attention_mask'] is None else torch.tensor([req['attention_mask']], device=dev)
        out = model.generate(input_ids=torch.tensor([req['prompt']], device=dev), attention_mask=am,
                             max_new_tokens=req['max_new_tokens'], eos_token_id=eos,
                             repetition_penalty=gen.get('repetition_penalty', 1.0), decoding_kwargs=dk,
                             return_dict_in_generate=True)
        assert out.sequences[0].tolist() == req['sequences'], (name, ri)
        assert out.kwargs['dls'] == req['dls'] and out.kwargs['edls'] == req['edls'], (name, ri)
        ids_log, n_log, mask_log = model.ids_log.cpu(), model.n_log.cpu(), model.mask_log.cpu().numpy().view(np.uint64)
        for k, st in enumerate(req['steps'][1:]):        # the draft the GPU trie produced == the reference's draft
            n = len(st['decoding_ids'])
            assert int(n_log[k]) == n and ids_log[k, :n].tolist() == st['decoding_ids'], (name, ri, k)
            assert np.array_equal(mask_log[k, :n], G.step_mask(st)), (name, ri, k)
        multi += sum(e > 1 for e in req['edls'])
    assert multi >= 3
