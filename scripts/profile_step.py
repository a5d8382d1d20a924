# -*- coding: utf-8 -*-
"""A short, profiler-friendly run of the hot path: Llama-2-7B shape, one prompt, a handful of decode steps.
Used under `ncu` (launch list / --set full); numbers printed here are never bench values."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--model', default='llama2-7b')
ap.add_argument('--new', type=int, default=12)
ap.add_argument('--prompt', type=int, default=256)
ap.add_argument('--requests', type=int, default=2)
ap.add_argument('--batch', type=int, default=0, help='> 0: the batched loop with this many requests')
ap.add_argument('--share', default='rows')
a = ap.parse_args()
from painlessinferenceacceleration_b200.common.lookahead_cache import LookaheadCache  # noqa: E402
from painlessinferenceacceleration_b200.models.llama.modeling_llama import LlamaForCausalLM  # noqa: E402

dev = torch.device('cuda:0')
cfg, _ = bench.make_config(a.model)
if a.batch:
    from painlessinferenceacceleration_b200.models.llama.modeling_llama_batch import LlamaForCausalLM  # noqa: E402,F811
model = LlamaForCausalLM(cfg, device=dev)
bench.synth_fill(model, cfg)
model.lookahead_cache = LookaheadCache(eos_ids=[2], device=dev, vocab_capacity=cfg.vocab_size, n_input_slots=max(a.batch, 8))
ps = bench.phrase_bank_prompts(max(4, a.batch), cfg.vocab_size, length=a.prompt)
prof = torch.cuda.cudart()   # run under `ncu --profile-from-start off`: only the LAST request is captured (the weight
                             # synthesis alone is thousands of elementwise launches)
if a.batch:
    for r in range(a.requests):
        if r == a.requests - 1:
            prof.cudaProfilerStart()
        o = model.generate(input_ids=torch.tensor(ps[:a.batch], device=dev), max_new_tokens=a.new, eos_token_id=2,
                           decoding_kwargs={'use_lookahead': True, 'decoding_length': 64, 'branch_length': 8,
                                            'batch_share': a.share}, return_dict_in_generate=True)
        print('batch request', r, 'edls', o.kwargs['edls'][:24], 'dls', o.kwargs['dls'][:24])
    prof.cudaProfilerStop()
    sys.exit(0)
for r in range(a.requests):  # same prompt twice: the second request drafts from the first one's answer
    if r == a.requests - 1:
        prof.cudaProfilerStart()
    o = model.generate(input_ids=torch.tensor([ps[0]], device=dev), max_new_tokens=a.new, eos_token_id=2,
                       decoding_kwargs={'use_lookahead': True, 'decoding_length': 64, 'branch_length': 8},
                       return_dict_in_generate=True)
    print('request', r, 'edls', o.kwargs['edls'], 'dls', o.kwargs['dls'])
prof.cudaProfilerStop()
