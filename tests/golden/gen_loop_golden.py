# -*- coding: utf-8 -*-
"""Generates tests/golden/loop_*.npz by running the REFERENCE'S OWN loop code (tests/golden/ref_loop.py: the
unmodified functions of /root/reference/lookahead/lookahead/common/pretrained_model.py:666-1268 + the live reference
trie) around installed tiny Hugging Face models.  Run in the build container only:

    python tests/golden/gen_loop_golden.py

Each fixture holds, per request, the prompt / parameters / resulting sequence / dls / edls, and per verify step
what the reference's accept routine saw (context ids, draft ids, tree mask rows, logits) and what it decided
(accepted tokens, kv_idx, continuous).  They pin oracle/loop.py (tests/test_loop_golden.py, CPU) and the CUDA
accept / compaction / device loop (tests/test_gpu_loop_golden.py) without the reference at run time.
Logits are stored in the model's dtype (bf16 as uint16 bit patterns), so that the repetition-penalty arithmetic of
the installed transformers' RepetitionPenaltyLogitsProcessor is reproduced bit for bit."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden.ref_loop import (import_reference, make_batch_driver, make_driver, run_reference_batch,  # noqa: E402
                                   run_reference_request)
from tests.tiny_models import prompts, tiny_hf_model  # noqa: E402


def mask_rows(m):
    m = np.asarray(m)
    return [int(sum(int(v) << j for j, v in enumerate(row.astype(np.int64).tolist()))) for row in m]


def logits_bits(t):
    t = t[0].contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    return t.float().numpy().copy()


def scenario(name, family, dtype, seed, vocab, requests, dl=64, bl=8, reps=2, **gen):
    """requests: list of dict(prompt=LongTensor[1,L], max_new_tokens=, attention_mask=None|tensor)"""
    _pm, _pmb, LookaheadCache = import_reference()
    torch.set_num_threads(4)
    hf = tiny_hf_model(family, seed=seed, dtype=dtype, vocab=vocab)
    rec = []
    trie = LookaheadCache()
    drv = make_driver(hf, trie, rec)
    reqs, arrays = [], {}
    for rep in range(reps):
        for q in requests:
            s0 = len(rec)
            r = run_reference_request(drv, q['prompt'], q['max_new_tokens'], attention_mask=q.get('attention_mask'),
                                      decoding_length=dl, branch_length=bl, **gen)
            steps = []
            for si in range(s0, len(rec)):
                st = rec[si]
                key = f'logits_{si}'
                arrays[key] = logits_bits(st['logits'])
                steps.append(dict(context_len=len(st['context']), decoding_ids=[int(x) for x in st['decoding_ids']],
                                  mask=None if st['decoding_masks'] is None else [str(v) for v in mask_rows(st['decoding_masks'])],
                                  logits=key, tokens=[int(x) for x in st['tokens']], dl=int(st['dl']), edl=int(st['edl']),
                                  kv=st['kv']))
            am = q.get('attention_mask')
            reqs.append(dict(prompt=q['prompt'][0].tolist(), max_new_tokens=q['max_new_tokens'],
                             attention_mask=None if am is None else am[0].tolist(), sequences=r['sequences'],
                             dls=r['dls'], edls=r['edls'], steps=steps))
    meta = dict(name=name, family=family, dtype=str(dtype).split('.')[-1], model_seed=seed, vocab=vocab,
                decoding_length=dl, branch_length=bl, gen={k: (sorted(v) if isinstance(v, set) else v) for k, v in gen.items()},
                requests=reqs)
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'loop_{name}.npz')
    np.savez_compressed(path, **arrays)
    edl = [e for r in reqs for e in r['edls'][1:]]
    print(f'{name}: {len(reqs)} requests, {len(rec)} steps, mean edl {np.mean(edl):.2f}, max edl {max(edl)}, '
          f'{os.path.getsize(path) // 1024} KB')


def batch_scenario(name, family, dtype, seed, vocab, batches, dl=64, bl=8, reps=2, **gen):
    """batches: list of (LongTensor [bs, L], max_new_tokens).  Runs the reference's BATCHED loop
    (pretrained_model_batch.py:1002-1330 + bat_get) and records every verify step."""
    _pm, _pmb, LookaheadCache = import_reference()
    torch.set_num_threads(4)
    hf = tiny_hf_model(family, seed=seed, dtype=dtype, vocab=vocab)
    rec = []
    drv = make_batch_driver(hf, LookaheadCache(), rec)
    calls, arrays = [], {}
    for rep in range(reps):
        for ids, mnt in batches:
            s0 = len(rec)
            r = run_reference_batch(drv, ids, mnt, decoding_length=dl, branch_length=bl, **gen)
            steps = []
            for si in range(s0, len(rec)):
                st = rec[si]
                key = f'logits_{si}'
                lg = st['logits'].contiguous()
                arrays[key] = lg.view(torch.int16).numpy().view(np.uint16).copy() if lg.dtype == torch.bfloat16 \
                    else lg.float().numpy().copy()
                steps.append(dict(prefill=st['prefill'], before=st['before'], logits=key, tokens=st['tokens'],
                                  dls=st['dls'], edls=st['edls']))
            calls.append(dict(input_ids=ids.tolist(), max_new_tokens=mnt, sequences=r['sequences'], dls=r['dls'],
                              edls=r['edls'], steps=steps))
    meta = dict(name=name, family=family, dtype=str(dtype).split('.')[-1], model_seed=seed, vocab=vocab,
                decoding_length=dl, branch_length=bl, gen=gen, calls=calls)
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'batchloop_{name}.npz')
    np.savez_compressed(path, **arrays)
    edl = [e for c in calls for e in c['edls'][len(c['input_ids']):]]
    print(f'{name}: {len(calls)} batches, {len(rec)} steps, mean edl {np.mean(edl):.2f}, max edl {max(edl)}, '
          f'{os.path.getsize(path) // 1024} KB')


def main():
    V = 96
    b3 = torch.cat(prompts(31, 3, 20, V), 0)
    b2 = torch.cat(prompts(32, 2, 16, V), 0)
    batch_scenario('llama_bf16_bs3_bs2', 'llama', torch.bfloat16, 2, V, [(b3, 32), (b2, 28)], pad_token_id=0)
    # an eos that one request meets early (requests leave the batch at different steps, _early_stop :937-980) +
    # repetition penalty
    hf = tiny_hf_model('mistral', seed=3, dtype=torch.bfloat16, vocab=V)
    _pm, _pmb, LookaheadCache = import_reference()
    b4 = torch.cat(prompts(33, 4, 14, V), 0)
    probe = run_reference_batch(make_batch_driver(hf, LookaheadCache(), None), b4, 30, repetition_penalty=1.1)
    eos = probe['sequences'][1][14 + 8]
    batch_scenario('mistral_bf16_bs4_rp11_eos', 'mistral', torch.bfloat16, 3, V, [(b4, 30)], repetition_penalty=1.1,
                   eos_token_id=int(eos), pad_token_id=0)
    ps = [dict(prompt=p, max_new_tokens=40) for p in prompts(21, 3, 24, V)]
    scenario('llama_bf16', 'llama', torch.bfloat16, 2, V, ps)
    scenario('mistral_bf16_rp11', 'mistral', torch.bfloat16, 3, V,
             [dict(prompt=p, max_new_tokens=36) for p in prompts(22, 3, 20, V)], repetition_penalty=1.1)
    # left padding (2-D attention mask, :1123-1131), an eos that is generated early (stop rule :1228-1231), stop words
    # (:1089, lookahead_cache.py:388-389,422-423) and a max_length that clamps branch_length near the end (:680)
    pad = 5
    reqs = []
    for p, mnt in zip(prompts(23, 3, 18, V), (30, 12, 7)):
        padded = torch.cat([torch.zeros((1, pad), dtype=torch.long), p], dim=1)
        am = torch.cat([torch.zeros((1, pad), dtype=torch.long), torch.ones_like(p)], dim=1)
        reqs.append(dict(prompt=padded, max_new_tokens=mnt, attention_mask=am))
    hf = tiny_hf_model('llama', seed=5, dtype=torch.bfloat16, vocab=V)
    # choose as eos a token the model generates a few steps into the first request
    _pm, _pmb, LookaheadCache = import_reference()
    drv = make_driver(hf, LookaheadCache(), None)
    probe = run_reference_request(drv, reqs[0]['prompt'], 30, attention_mask=reqs[0]['attention_mask'])
    eos = probe['sequences'][reqs[0]['prompt'].shape[1] + 9]
    sw = {probe['sequences'][reqs[0]['prompt'].shape[1] + 3]}
    scenario('llama_bf16_pad_eos_stop', 'llama', torch.bfloat16, 5, V, reqs, eos_token_id=int(eos), stop_words=sw)
    # BASELINE config 1: GPT-2 shape, fp32, 16-token / 4-branch drafts
    scenario('gpt2_fp32_16_4', 'gpt2', torch.float32, 1, V,
             [dict(prompt=p, max_new_tokens=32) for p in prompts(24, 3, 12, V)], dl=16, bl=4)
    # other draft formats inside the loop (decoding_mode, :712-715): one branch.  `par` cannot be recorded: par_get
    # returns a float64 mask (lookahead_cache.py:481 np.tril(np.ones(..))), so the reference's own accept routine dies
    # with "TypeError: slice indices must be integers" at pretrained_model.py:819 on the first non-empty draft.
    scenario('llama_bf16_one', 'llama', torch.bfloat16, 2, V, ps[:2], decoding_mode='one')


if __name__ == '__main__':
    main()
