# -*- coding: utf-8 -*-
"""CPU/torch restatement of the reference's bs=1 draft -> verify -> accept loop -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this.

The reference loop (/root/reference/lookahead/lookahead/common/pretrained_model.py) cannot be imported under the
installed transformers 5.5 (it pins 4.30.2 / 4.36.0), so it is restated here on top of the INSTALLED
Hugging Face models driven with the reference's patch semantics (rank-4 0/1 mask -> additive mask + position ids
= rowsum - 1, models/llama/modeling_llama.py:584-588):

    draft + mask      pretrained_model.py:666-756     -> _draft()
    accept walk       pretrained_model.py:764-892     -> _accept()
    KV compaction     pretrained_model.py:894-945     -> _compact()
    loop / trie puts  pretrained_model.py:1085-1239   -> lookahead_generate()

"parity unpinned" for the logits-processor arithmetic: it is the installed transformers'
RepetitionPenaltyLogitsProcessor (the reference only calls it, :786/:834, and ships no test for it).
The trie is any object with the LookaheadCache surface (oracle.trie.OracleLookaheadCache in tests)."""
import time

import numpy as np
import torch


def _additive(mask01, dtype):
    # (1 - mask) * finfo.min  (modeling_llama.py:588)
    return (1.0 - mask01.to(dtype)) * torch.finfo(dtype).min


def _cache_rows(cache):
    return cache.layers[0].keys.shape[2] if len(cache.layers) and cache.layers[0].keys is not None else 0


def _compact(cache, keep_idx):
    """_update_cache_with_axis_2 (:894-907): keep rows [0, context) + kv_idx on the sequence axis"""
    for layer in cache.layers:
        layer.keys = layer.keys[:, :, keep_idx]
        layer.values = layer.values[:, :, keep_idx]


def _penalise(scores, input_ids, penalty):
    if penalty == 1.0:
        return scores
    from transformers import RepetitionPenaltyLogitsProcessor
    return RepetitionPenaltyLogitsProcessor(penalty)(input_ids, scores)


def _accept(decoding_ids, decoding_masks, logits, input_ids, penalty):
    """_lookahead_update_model_kwargs_for_generation (:764-892), greedy. returns (tokens, logit_indices)"""
    draft_ids = decoding_ids[1:]
    update_ids = input_ids
    if len(draft_ids) == 0:  # :783-798
        scores = _penalise(logits[:, -1], update_ids, penalty)
        return [int(torch.argmax(scores, dim=-1))], [0]
    dm = decoding_masks[1:, 1:]
    depth = dm.sum(axis=1)
    max_depth = int(depth.max())
    # leaves: nodes whose successor in DFS order is not deeper, plus the last node (:806-814)
    leaves = [i - 1 for i in range(1, len(depth)) if depth[i] <= depth[i - 1]] + [len(depth) - 1]
    paths = [np.flatnonzero(dm[leaf]).tolist() for leaf in leaves]          # root-to-leaf node indices
    branches = [[draft_ids[i] for i in p] for p in paths]
    tokens, logit_indices = [], []
    for i in range(-1, max_depth):  # :827-860
        logit_index = 0 if i == -1 else paths[0][i] + 1
        scores = _penalise(logits[:, logit_index], update_ids, penalty)
        nxt = int(torch.argmax(scores, dim=-1))
        update_ids = torch.cat([update_ids, torch.tensor([[nxt]], dtype=update_ids.dtype, device=update_ids.device)], 1)
        tokens.append(nxt)
        logit_indices.append(logit_index)
        if i == max_depth - 1:
            break
        keep = [k for k, b in enumerate(branches) if len(b) > i + 1 and b[i + 1] == nxt]
        if not keep:
            break
        paths = [paths[k] for k in keep]
        branches = [branches[k] for k in keep]
    return tokens, logit_indices


class HFBackend(object):
    """verify forward + KV cache of an installed Hugging Face causal LM (eager attention)"""

    def __init__(self, model):
        from transformers import DynamicCache
        self.model = model
        self.cache = DynamicCache(config=model.config)
        self.dtype = next(model.parameters()).dtype

    def rows(self):
        return _cache_rows(self.cache)

    def forward(self, ids_in, m01, pos):
        return self.model(input_ids=ids_in, attention_mask=_additive(m01, self.dtype), position_ids=pos,
                          past_key_values=self.cache, use_cache=True).logits

    def compact(self, keep_idx):
        _compact(self.cache, keep_idx)


@torch.no_grad()
def lookahead_generate(model, trie, input_ids, max_new_tokens=None, max_length=None, eos_token_id=(2,),
                       decoding_length=64, branch_length=8, decoding_mode='hier', max_query_length=2,
                       repetition_penalty=1.0, use_lookahead=True, stop_words=None, attention_mask=None,
                       trace=False, time_budget_s=None, backend=None, min_forwards=2):
    """restates lookahead_generation (:947-1268) for one request. input_ids: LongTensor [1, len].
    returns dict(sequences, dls, edls, fts, qts[, steps])"""
    assert input_ids.size(0) == 1
    dev = input_ids.device
    be = backend if backend is not None else HFBackend(model)
    prompt_len = input_ids.shape[1]
    if max_length is None:
        max_length = prompt_len + int(max_new_tokens)
    eos = list(eos_token_id) if eos_token_id is not None else []
    if use_lookahead:
        trie.eos_ids = eos if eos else None
        trie.stop_words = stop_words if stop_words is not None else {}
    Lmax = max_length + decoding_length + 1                                   # :1115
    pad = torch.ones((1, Lmax), dtype=torch.long, device=dev)
    if attention_mask is not None:
        pad[0, :attention_mask.shape[1]] = attention_mask[0].long()
    full = torch.tril(pad[:, None, None].expand(-1, -1, Lmax, -1), 0)         # :1131
    dls, edls, fts, qts, steps = [], [], [], [], []
    if use_lookahead:
        trie.put(input_ids[0].tolist()[1:], branch_length=branch_length + 1, mode='input', idx=0)   # :1156
    ts = t_start = time.time()
    while True:
        cur_len = input_ids.shape[1]
        if be.rows() == 0:
            # prefill (:683-705): mask[:, :, :len, :len]
            m01 = full[:, :, :cur_len, :cur_len]
            ids_in = input_ids
            decoding_ids, decoding_masks = None, None
        else:
            ubl = min(branch_length, max_length - cur_len - 1)                # :680
            assert ubl >= 0
            qids = input_ids[0, -max_query_length:].tolist()                 # :708
            t0 = time.time()
            if use_lookahead:
                mode = decoding_mode if '_' in decoding_mode else decoding_mode + '_mix'
                fmt, tmode = mode.split('_')
                decoding_ids, decoding_masks, _sizes = getattr(trie, fmt + '_get')(
                    qids, decoding_length=decoding_length, branch_length=ubl, min_input_size=0,
                    min_output_size=max(decoding_length // 2, 1), mode=tmode, idx=0)      # :709-723
            else:
                decoding_ids, decoding_masks = qids[-1:], np.ones((1, 1), dtype=np.int64)
            qts.append(time.time() - t0)
            P, n = cur_len - 1, len(decoding_ids)
            ids_in = torch.tensor([decoding_ids], dtype=torch.long, device=dev)
            tree = torch.from_numpy(np.asarray(decoding_masks)[None, None]).to(device=dev, dtype=torch.long)
            m01 = torch.cat([full[:, :, P:P + n, :P], tree], dim=3)           # :731-734
        pos = (m01.sum(-1).squeeze(1) - 1).clamp(min=0)                      # modeling_llama.py:587
        logits = be.forward(ids_in, m01, pos)
        if decoding_ids is None:
            tokens, logit_indices = _accept([0], None, logits, input_ids, repetition_penalty)
            dls.append(1)
        else:
            tokens, logit_indices = _accept(decoding_ids, np.asarray(decoding_masks), logits, input_ids,
                                            repetition_penalty)
            n_draft = len(decoding_ids) - 1
            if n_draft != len(tokens) - 1:                                    # :865-875
                kv_idx = [li - 1 + cur_len for li in logit_indices[1:]]
                keep = torch.tensor(list(range(cur_len)) + kv_idx, dtype=torch.long, device=dev)
                be.compact(keep)
            dls.append(n_draft + 1)
        edls.append(len(tokens))
        if trace:
            steps.append(dict(decoding_ids=list(decoding_ids) if decoding_ids is not None else None,
                              tokens=list(tokens), logit_indices=list(logit_indices)))
        input_ids = torch.cat([input_ids, torch.tensor([tokens], dtype=torch.long, device=dev)], dim=1)
        if use_lookahead:
            trie.stream_put(tokens, branch_length=branch_length + 1, final=False, mode='output', idx=0)   # :1203
        done = input_ids.shape[1] >= max_length or any(e in tokens for e in eos)                         # :1225-1231
        te = time.time()
        if time_budget_s is not None and te - t_start > time_budget_s and len(edls) >= min_forwards:  # bounded sample
            done = True
        fts.append(te - ts)
        ts = te
        if done:
            if use_lookahead:
                trie.stream_put([], branch_length=branch_length + 1, final=True, mode='output', idx=0)    # :1237
            break
    res = dict(sequences=input_ids, dls=dls, edls=edls, fts=fts, qts=qts)
    if trace:
        res['steps'] = steps
    return res


@torch.no_grad()
def greedy_generate(model, input_ids, max_new_tokens, eos_token_id=(2,), repetition_penalty=1.0):
    """plain greedy decoding of the same HF model (the lossless-property partner, lookahead/README.md:45)"""
    return lookahead_generate(model, None, input_ids, max_new_tokens=max_new_tokens, eos_token_id=eos_token_id,
                              decoding_length=1, branch_length=1, repetition_penalty=repetition_penalty,
                              use_lookahead=False)
