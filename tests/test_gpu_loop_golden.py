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
            acc.run(logits, ids, mask, dn, seq, seq_len, toks, cnt, nodes, prefix, fin)
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
    recorded_moves = sum(1 for r in meta['requests'] for st in r['steps'] if st['kv'] is not None and not st['kv']['continuous'])
    assert n_checked >= 15 and n_moved == recorded_moves


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
            assert len(steps) < self.cap  # + the step launched ahead of the stop check
            self.all.zero_()
            for k, st in enumerate(steps[1:]):
                self.all[k] = _pad_logits(meta, z, st).to(dev)
            self.first[0] = G.step_logits(meta, z, steps[0])[0, -1].to(dev)
            self.step.zero_()
            self.ids_log.zero_()

        def _prefill_logits(self, rt, prompt_len, slot=0, row=0):
            rt.logits[row:row + 1] = self.first

        def _verify_layers(self, rt, bufs=None, last_only=False):
            rt.logits.copy_(self.all.index_select(0, self.step)[0])
            self.ids_log.index_copy_(0, self.step, rt.ids[None])
            self.n_log.index_copy_(0, self.step, rt.n)
            self.mask_log.index_copy_(0, self.step, rt.mask[None, :, 0])
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


# ---------------------------------------------------------------------------------------------------------------
# batched device loop (common/pretrained_model_batch.py) against records of the reference's own batched loop
# ---------------------------------------------------------------------------------------------------------------
def _batch_replay_model(vocab, dev, dl):
    from painlessinferenceacceleration_b200.common.pretrained_model_batch import LookaheadPreTrainedModel

    class BatchReplay(LookaheadPreTrainedModel):
        def __init__(self):
            super().__init__(types.SimpleNamespace(eos_token_id=2, pad_token_id=0))
            self.anchor = nn.Parameter(torch.zeros(1, device=dev), requires_grad=False)
            self.cap = 64
            self.all = torch.zeros((self.cap, 64, vocab), dtype=torch.bfloat16, device=dev)
            self.ids_log = torch.zeros((self.cap, 64), dtype=torch.int32, device=dev)
            self.n_log = torch.zeros((self.cap, 8), dtype=torch.int32, device=dev)
            self.step = torch.zeros((1,), dtype=torch.int64, device=dev)
            self.first = None

        def geometry(self):
            return dict(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, head_dim=128, inter=128, vocab=vocab)

        def rope_tables(self, max_pos):
            z = torch.zeros((max_pos, 64), dtype=torch.bfloat16, device=dev)
            return z, z.clone()

        @staticmethod
        def share(k):
            return max(dl // k, 1) // k   # decoding_length // active // active (pretrained_model_batch.py:713 + bat_get :534)

        def load(self, meta, z, call):
            steps = call['steps']
            assert steps[0]['prefill'] and len(steps) < self.cap
            self.first = G.batch_step_logits(meta, z, steps[0])[:, -1].to(dev)       # [bs, V]
            self.all.zero_()
            for t, st in enumerate(steps[1:]):
                lg = G.batch_step_logits(meta, z, st)
                k, n = lg.shape[0], lg.shape[1]
                sh = self.share(k)
                assert n <= sh
                for r in range(k):
                    self.all[t, r * sh:r * sh + n] = lg[r].to(dev)
            self.step.zero_()
            self.ids_log.zero_()
            self.n_log.zero_()

        def _prefill_logits(self, rt, prompt_len, slot=0, row=0):
            rt.logits[row:row + 1] = self.first[slot:slot + 1]

        def _verify_layers(self, rt, bufs=None, last_only=False):
            rt.logits.copy_(self.all.index_select(0, self.step)[0])
            self.ids_log.index_copy_(0, self.step, rt.ids[None])
            self.n_log[:, :rt.n.numel()].index_copy_(0, self.step, rt.n[None])
            self.step += 1

    return BatchReplay()


@pytest.mark.parametrize('name', G.batch_names())
def test_batched_device_loop_reproduces_the_reference_batch_loop(name):
    """GPU trie batched get (request idx per row) -> [recorded logits] -> per-slot accept (bounded walk) -> per-slot
    KV compaction -> per-slot stream_put with device-resident idx, slots compacted as requests finish: the drafts,
    tokens, dls, edls and the padded output equal the reference's batched loop for every call"""
    from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache
    meta, z = G.load_batch(name)
    gen = meta['gen']
    eos, V, dl = gen.get('eos_token_id', 2), meta['vocab'], meta['decoding_length']
    dev = torch.device(DEV)
    model = _batch_replay_model(V, dev, dl)
    model.lookahead_cache = LookaheadCache(eos_ids=[eos], device=dev, vocab_capacity=1024, node_capacity=1 << 20)
    multi = 0
    for ci, call in enumerate(meta['calls']):
        model.load(meta, z, call)
        ids = torch.tensor(call['input_ids'], device=dev)
        out = model.generate(input_ids=ids, max_new_tokens=call['max_new_tokens'], eos_token_id=eos,
                             pad_token_id=gen.get('pad_token_id', 0),
                             repetition_penalty=gen.get('repetition_penalty', 1.0),
                             decoding_kwargs={'use_lookahead': True, 'decoding_length': dl,
                                              'branch_length': meta['branch_length']},
                             return_dict_in_generate=True)
        assert out.sequences.tolist() == call['sequences'], (name, ci)
        assert out.kwargs['dls'] == call['dls'] and out.kwargs['edls'] == call['edls'], (name, ci)
        ids_log, n_log = model.ids_log.cpu(), model.n_log.cpu()
        for t, st in enumerate(call['steps'][1:]):
            k = len(st['before']['batch_indices'])
            sh = model.share(k)
            for r in range(k):
                n = int(n_log[t, r])
                want = st['before']['ids'][r]
                assert ids_log[t, r * sh:r * sh + n].tolist() == want[:n] and all(x == 0 for x in want[n:]), (name, ci, t, r)
        multi += sum(e > 1 for e in call['edls'])
    assert multi >= 3
