# -*- coding: utf-8 -*-
"""Llama with the BATCHED lookahead loop (reference: models/llama/modeling_llama_batch.py - the batch variant of the
patched model: preallocated KV cache addressed by per-request cursors :375-405, fused RoPE :188-189).  The module
tree, weights and verify kernels are those of modeling_llama.py; the request-slot runtime gives every request its own
cursor-addressed cache, and common/pretrained_model_batch.py drives all requests through one verify forward."""
from ...common.pretrained_model_batch import LookaheadPreTrainedModel as _BatchLoop
from .modeling_llama import LlamaForCausalLM as _LlamaForCausalLM


class LlamaForCausalLM(_BatchLoop, _LlamaForCausalLM):
    pass
