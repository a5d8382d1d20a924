# -*- coding: utf-8 -*-
"""Mistral with the lookahead patch (reference: models/mistral/modeling_mistral.py, patch :979-982, attention
:241-320, repeat_kv :183-192, rope :94-164).  Same decoder as Llama with grouped-query attention: the tree
attention kernel packs two query heads of one KV head into each UMMA M=128 tile (no repeat_kv copy), and - as on
the reference's lookahead branch - the sliding window is ignored (:979-982 vs :1016-1022)."""
from ..llama.modeling_llama import LlamaForCausalLM


class MistralForCausalLM(LlamaForCausalLM):
    pass
