# -*- coding: utf-8 -*-
"""Picks bench.py's EMBED_STD: for a few embedding scales, how often does greedy decoding of the synthetic model follow
the successor chain, and what accepted length does a trie warmed on OTHER prompts reach (first pass)?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache  # noqa: E402
from painlessinferenceacceleration_b200.models.llama.modeling_llama import LlamaForCausalLM  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'llama2-7b'
scales = [float(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0.25, 0.5, 1.0, 2.0, 4.0]
dev = torch.device('cuda:0')
cfg, fam = bench.make_config(name)
model = LlamaForCausalLM(cfg, device=dev)
succ = bench.successor_map(cfg.vocab_size).tolist()
allp = bench.phrase_bank_prompts(72, cfg.vocab_size)
for es in scales:
    bench.synth_fill(model, cfg, embed_std=es)
    model._tiled_weights = {}
    model._rt = None
    model._fused = False
    model.lookahead_cache = LookaheadCache(eos_ids=[2], device=dev, vocab_capacity=cfg.vocab_size)
    dk = {'use_lookahead': True, 'decoding_length': 64, 'branch_length': 8}
    rows = []
    for i, p in enumerate([allp[64 + j] for j in range(4)] + allp[:6]):
        o = model.generate(input_ids=torch.tensor([p], device=dev), max_new_tokens=256, eos_token_id=2,
                           decoding_kwargs=dict(dk), return_dict_in_generate=True)
        seq = o.sequences[0].tolist()
        gen = seq[256:]
        follow = sum(1 for a, b in zip(seq[255:-1], gen) if succ[a] == b) / max(len(gen), 1)
        rows.append((float(np.mean(o.kwargs['edls'][1:])), follow, len(set(gen)) / max(len(gen), 1)))
    print(f'embed_std {es}: edl per request {[round(r[0], 2) for r in rows]}  follow-succ {[round(r[1], 2) for r in rows]} '
          f'distinct-token ratio {[round(r[2], 2) for r in rows]}', flush=True)
