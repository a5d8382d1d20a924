# -*- coding: utf-8 -*-
"""The CPU restatement of the loop (oracle/loop.py) against the one result the reference pins for it: the
lossless property -- lookahead output == plain greedy output of the same model in fp32
(lookahead/README.md:45; examples/llama_example.py:39-69 run both).  Covers BASELINE config 1 (GPT-2 shape,
16-token / 4-branch drafts on CPU) and the Llama / Mistral / Mixtral patch surfaces."""
import pytest
import torch

from oracle.loop import greedy_generate, lookahead_generate
from oracle.trie import OracleLookaheadCache
from tests.tiny_models import prompts, tiny_hf_model


@pytest.mark.parametrize('family,dl,bl', [('gpt2', 16, 4), ('llama', 64, 8), ('mistral', 64, 8), ('mixtral', 32, 6)])
def test_lossless_fp32(family, dl, bl):
    torch.set_num_threads(4)
    model = tiny_hf_model(family, seed=1)
    trie = OracleLookaheadCache(eos_ids=[2])
    total_edl = []
    for rep in range(2):
        for k, p in enumerate(prompts(7, 3, 12, 64)):
            ref = greedy_generate(model, p, max_new_tokens=40)
            out = lookahead_generate(model, trie, p, max_new_tokens=40, decoding_length=dl, branch_length=bl)
            assert out['sequences'].tolist() == ref['sequences'].tolist(), f'{family} prompt {k}'
            assert out['sequences'].shape[1] <= 12 + 40
            assert sum(out['edls']) == out['sequences'].shape[1] - 12
            if rep == 1:
                total_edl += out['edls'][1:]
    # the trie is carried across requests: the second pass drafts the first pass's answers
    assert sum(total_edl) / len(total_edl) > 2.0, total_edl


def test_repetition_penalty_lossless():
    model = tiny_hf_model('mistral', seed=3)
    trie = OracleLookaheadCache(eos_ids=[2])
    for p in prompts(9, 3, 10, 64):
        ref = greedy_generate(model, p, max_new_tokens=32, repetition_penalty=1.1)
        out = lookahead_generate(model, trie, p, max_new_tokens=32, repetition_penalty=1.1)
        assert out['sequences'].tolist() == ref['sequences'].tolist()


def test_left_padding_and_max_length():
    model = tiny_hf_model('llama', seed=5)
    trie = OracleLookaheadCache(eos_ids=[2])
    p = prompts(11, 1, 9, 64)[0]
    padded = torch.cat([torch.zeros((1, 3), dtype=torch.long), p], dim=1)
    am = torch.cat([torch.zeros((1, 3), dtype=torch.long), torch.ones((1, 9), dtype=torch.long)], dim=1)
    a = lookahead_generate(model, trie, padded, max_new_tokens=20, attention_mask=am)
    b = lookahead_generate(model, trie, padded, max_new_tokens=20, attention_mask=am)
    assert a['sequences'].tolist() == b['sequences'].tolist()
    assert a['sequences'].shape[1] <= 12 + 20
    assert max(b['edls']) > 1  # second run drafts the first run's output from the trie
