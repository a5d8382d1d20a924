# -*- coding: utf-8 -*-
"""Runs the reference's OWN generation-loop code (bs = 1 and batched) in this container -- fixture generator
infrastructure, never imported by tests at run time on the GPU box (there is no /root/reference there).

The reference loop (/root/reference/lookahead/lookahead/common/pretrained_model.py and pretrained_model_batch.py)
pins transformers 4.30.2 / 4.36.0 and fails to import under the installed 5.5 because of four names that its
lookahead path never touches (beam-search scorers/constraints and three output dataclasses).  With
those four names stubbed the module imports, and its methods
    lookahead_generation                              :947-1268
    lookahead_prepare_inputs_for_generation           :666-756
    _lookahead_update_model_kwargs_for_generation     :764-892
    _update_cache / _update_cache_with_axis_2         :894-945
are plain functions: they are borrowed, unmodified, by `RefDriver`, whose only own code is what the reference
expects from its host class: `self(...)` (= the patched model forward: rank-4 0/1 mask -> position_ids = rowsum - 1
and additive mask, models/llama/modeling_llama.py:584-588, evaluated by an INSTALLED Hugging Face model),
`_extract_past_from_model_output`, `config`, `generation_config`.  The trie is the live reference LookaheadCache.
Nothing of the reference is copied into the repository: the code is imported from where it lies."""
import sys
import types

import numpy as np
import torch

REF = '/root/reference/lookahead'


def import_reference():
    """the reference's pretrained_model / pretrained_model_batch modules under the installed transformers"""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import transformers.generation.utils as gu
    for m, names in [('transformers.generation.beam_constraints', ['DisjunctiveConstraint', 'PhrasalConstraint']),
                     ('transformers.generation.beam_search', ['BeamSearchScorer', 'ConstrainedBeamSearchScorer'])]:
        if m not in sys.modules:
            mod = types.ModuleType(m)
            for n in names:
                setattr(mod, n, type(n, (), {}))
            sys.modules[m] = mod
    for n in ['GreedySearchEncoderDecoderOutput', 'GreedySearchDecoderOnlyOutput', 'GreedySearchOutput', 'SampleOutput']:
        if not hasattr(gu, n):
            setattr(gu, n, type(n, (), {}))
    import lookahead.common.pretrained_model as pm
    import lookahead.common.pretrained_model_batch as pmb
    from lookahead.common.lookahead_cache import LookaheadCache
    return pm, pmb, LookaheadCache


class _Out(dict):
    __getattr__ = dict.get


def _hf_forward(hf, input_ids, attention_mask, past):
    """the reference's patched forward (modeling_llama.py:584-588) on an installed HF model: legacy (k, v) tuples in
    and out, positions = mask row sums - 1, additive mask (1 - m) * finfo.min"""
    from transformers import DynamicCache
    dtype = next(hf.parameters()).dtype
    cache = DynamicCache(config=hf.config)
    if past is not None:
        for li, (k, v) in enumerate(past):
            cache.update(k, v, li)
    pos = (attention_mask.sum(-1).squeeze(1) - 1).clamp(min=0)
    add = (1.0 - attention_mask.to(dtype)) * torch.finfo(dtype).min
    out = hf(input_ids=input_ids, attention_mask=add, position_ids=pos, past_key_values=cache, use_cache=True)
    new_past = tuple((layer.keys, layer.values) for layer in cache.layers)
    return out.logits, new_past


def make_driver(hf, trie=None, record=None):
    """an object carrying the reference's loop methods around an installed HF causal LM.  `record` (a list) receives
    one dict per verify step: what the accept routine saw and what it decided."""
    pm, _pmb, LookaheadCache = import_reference()
    P = pm.LookaheadPreTrainedModel

    class RefDriver(object):
        lookahead_generation = P.lookahead_generation
        lookahead_prepare_inputs_for_generation = P.lookahead_prepare_inputs_for_generation
        _ref_update = P._lookahead_update_model_kwargs_for_generation
        _update_cache_with_axis_2 = P._update_cache_with_axis_2
        _ref_update_cache = P._update_cache
        _get_position_ids = P._get_position_ids

        def __init__(self):
            self.config = hf.config
            self.generation_config = hf.generation_config
            self._kv = None

        def __call__(self, input_ids=None, past_key_values=None, use_cache=None, attention_mask=None, return_dict=True,
                     output_attentions=None, output_hidden_states=None, position_ids=None, **kw):
            logits, past = _hf_forward(hf, input_ids, attention_mask, past_key_values)
            return _Out(logits=logits, past_key_values=past)

        def _extract_past_from_model_output(self, outputs, standardize_cache_format=False):
            return outputs.past_key_values

        def _update_cache(self, past_key_values, kv_idx, context_length=None, max_match_count=None, continuous=False):
            self._kv = dict(kv_idx=kv_idx.tolist(), continuous=bool(continuous))
            return self._ref_update_cache(past_key_values, kv_idx, context_length=context_length,
                                          max_match_count=max_match_count, continuous=continuous)

        def _lookahead_update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False,
                                                          standardize_cache_format=False, logits_processor=None,
                                                          input_ids=None):
            dk = model_kwargs['decoding_kwargs']
            ids = list(dk.get('decoding_ids', []))
            masks = dk.get('decoding_masks', None)
            self._kv = None
            n_before = len(dk['edls'])
            res = self._ref_update(outputs, model_kwargs, is_encoder_decoder=is_encoder_decoder,
                                   standardize_cache_format=standardize_cache_format,
                                   logits_processor=logits_processor, input_ids=input_ids)
            if record is not None:
                ntl = res['next_token_list']
                record.append(dict(context=input_ids[0].tolist(), decoding_ids=ids,
                                   decoding_masks=None if masks is None else np.asarray(masks).copy(),
                                   logits=outputs.logits.detach().clone(),
                                   tokens=list(ntl[0]), dl=dk['dls'][n_before], edl=dk['edls'][n_before], kv=self._kv))
            return res

    d = RefDriver()
    if trie is not None:
        d.lookahead_cache = trie
    return d


def run_reference_request(driver, input_ids, max_new_tokens, eos_token_id=2, repetition_penalty=1.0,
                          attention_mask=None, decoding_length=64, branch_length=8, decoding_mode='hier',
                          stop_words=None, do_sample=False):
    """one call of the reference's lookahead_generation() exactly as its generate() would make it (:349-372, :652-664)"""
    from transformers import LogitsProcessorList, MaxLengthCriteria, RepetitionPenaltyLogitsProcessor, StoppingCriteriaList
    max_length = input_ids.shape[1] + max_new_tokens
    lp = LogitsProcessorList()
    if repetition_penalty != 1.0:
        lp.append(RepetitionPenaltyLogitsProcessor(penalty=repetition_penalty))
    sc = StoppingCriteriaList([MaxLengthCriteria(max_length=max_length)])
    dk = {'use_lookahead': True, 'decoding_length': decoding_length, 'branch_length': branch_length,
          'decoding_mode': decoding_mode, 'do_sample': do_sample}
    if stop_words is not None:
        dk['stop_words'] = stop_words
    kw = dict(decoding_kwargs=dk, use_cache=True)
    if attention_mask is not None:
        kw['attention_mask'] = attention_mask
    with torch.no_grad():
        out = driver.lookahead_generation(input_ids, logits_processor=lp, stopping_criteria=sc, pad_token_id=0,
                                          eos_token_id=eos_token_id, output_scores=False, return_dict_in_generate=True,
                                          output_attentions=False, output_hidden_states=False, **kw)
    return dict(sequences=out.sequences[0].tolist(), dls=list(out.kwargs['dls']), edls=list(out.kwargs['edls']))


# ----------------------------------------------------------------------------------------------------------------
# batched loop: /root/reference/lookahead/lookahead/common/pretrained_model_batch.py
#   lookahead_generation :1002-1330, lookahead_prepare_inputs_for_generation :664-759,
#   _lookahead_update_model_kwargs_for_generation :767-935, _early_stop :937-980, _update_cache :982-989
# borrowed unmodified; the driver's own code is the batched patched forward the reference expects from its model
# (models/llama/modeling_llama_batch.py:355-405: preallocated [bs, H, decoding_max_length, D] caches, request b's draft
# rows written at its cursor), evaluated request by request with an installed HF model.
# ----------------------------------------------------------------------------------------------------------------
def make_batch_driver(hf, trie, record=None):
    _pm, pmb, _LC = import_reference()
    P = pmb.LookaheadPreTrainedModel

    class RefBatchDriver(object):
        lookahead_generation = P.lookahead_generation
        lookahead_prepare_inputs_for_generation = P.lookahead_prepare_inputs_for_generation
        _ref_update = P._lookahead_update_model_kwargs_for_generation
        _early_stop = P._early_stop
        _update_cache = P._update_cache
        _update_cache_with_axis_2 = P._update_cache_with_axis_2
        _get_position_ids = P._get_position_ids

        def __init__(self):
            self.config = hf.config
            self.generation_config = hf.generation_config
            self.lookahead_cache = trie

        def __call__(self, input_ids=None, past_key_values=None, use_cache=None, attention_mask=None, return_dict=True,
                     output_attentions=None, output_hidden_states=None, position_ids=None, decoding_kwargs=None, **kw):
            bs, n = input_ids.shape
            if past_key_values is None:
                Lmax = decoding_kwargs['decoding_max_length']
                outs = [_hf_forward(hf, input_ids[b:b + 1], attention_mask[b:b + 1], None) for b in range(bs)]
                past = []
                for li in range(len(outs[0][1])):
                    k = torch.cat([o[1][li][0] for o in outs], 0)
                    v = torch.cat([o[1][li][1] for o in outs], 0)
                    z = torch.zeros((bs, k.shape[1], Lmax - n, k.shape[3]), dtype=k.dtype)
                    past.append((torch.cat([k, z], 2), torch.cat([v, z], 2)))
                return _Out(logits=torch.cat([o[0] for o in outs], 0), past_key_values=tuple(past))
            cursors = decoding_kwargs['decoding_cursors']
            logits = []
            for b in range(bs):
                cur = cursors[b]
                pb = tuple((k[b:b + 1, :, :cur], v[b:b + 1, :, :cur]) for k, v in past_key_values)
                lg, newp = _hf_forward(hf, input_ids[b:b + 1], attention_mask[b:b + 1, :, :, :cur + n], pb)
                for (k, v), (nk, nv) in zip(past_key_values, newp):
                    k[b, :, cur:cur + n] = nk[0, :, cur:cur + n]
                    v[b, :, cur:cur + n] = nv[0, :, cur:cur + n]
                logits.append(lg)
            return _Out(logits=torch.cat(logits, 0), past_key_values=past_key_values)

        def _lookahead_update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False,
                                                          standardize_cache_format=False, logits_processor=None,
                                                          input_ids=None):
            dk = model_kwargs['decoding_kwargs']
            prefill = model_kwargs.get('past_key_values', None) is None
            before = None if prefill else dict(ids=[list(x) for x in dk['decoding_ids']],
                                               cursors=list(dk['decoding_cursors']),
                                               batch_indices=list(dk['batch_indices']))
            n0 = len(dk['edls'])
            res = self._ref_update(outputs, model_kwargs, is_encoder_decoder=is_encoder_decoder,
                                   standardize_cache_format=standardize_cache_format,
                                   logits_processor=logits_processor, input_ids=input_ids)
            if record is not None:
                record.append(dict(prefill=prefill, before=before, logits=outputs.logits.detach().clone(),
                                   tokens=[list(t) for t in res['next_token_list']],
                                   dls=list(dk['dls'][n0:]), edls=list(dk['edls'][n0:])))
            return res

    return RefBatchDriver()


def run_reference_batch(driver, input_ids, max_new_tokens, eos_token_id=2, repetition_penalty=1.0,
                        decoding_length=64, branch_length=8, decoding_mode='hier', pad_token_id=0):
    """one call of the reference's batched lookahead_generation()"""
    from transformers import LogitsProcessorList, MaxLengthCriteria, RepetitionPenaltyLogitsProcessor, StoppingCriteriaList
    max_length = input_ids.shape[1] + max_new_tokens
    lp = LogitsProcessorList()
    if repetition_penalty != 1.0:
        lp.append(RepetitionPenaltyLogitsProcessor(penalty=repetition_penalty))
    sc = StoppingCriteriaList([MaxLengthCriteria(max_length=max_length)])
    dk = {'use_lookahead': True, 'decoding_length': decoding_length, 'branch_length': branch_length,
          'decoding_mode': decoding_mode, 'do_sample': False}
    with torch.no_grad():
        out = driver.lookahead_generation(input_ids.clone(), logits_processor=lp, stopping_criteria=sc,
                                          pad_token_id=pad_token_id, eos_token_id=eos_token_id, output_scores=False,
                                          return_dict_in_generate=True, output_attentions=False,
                                          output_hidden_states=False, decoding_kwargs=dk, use_cache=True)
    return dict(sequences=out.sequences.tolist(), dls=list(out.kwargs['dls']), edls=list(out.kwargs['edls']))
