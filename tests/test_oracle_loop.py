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


@pytest.mark.parametrize('family,mseed,pseed', [('llama', 3, 16), ('llama', 4, 17), ('mistral', 2, 15)])
def test_batched_loop_is_lossless_fp32(family, mseed, pseed):
    """oracle/loop_batch.py (pretrained_model_batch.py:664-1330 restated): every request of a batch yields
    exactly the tokens plain greedy decoding yields for it; per-slot trie input/output puts, bat_get drafts of
    decoding_length // active requests, requests leaving the batch as they finish (the seeds give requests of
    different lengths: some end on eos early, some run to max_length).
    Token id 0 must not occur in the answers: `ids[0] = match_token_id or self.token_id` (lookahead_cache.py:129)
    replaces a root token 0 by the tree's key token - a quirk of the reference that both restatements and the CUDA
    trie keep, and that breaks the lossless property for that (in real vocabularies never generated) id."""
    from oracle.loop_batch import lookahead_generate_batch
    torch.set_num_threads(4)
    model = tiny_hf_model(family, seed=mseed)
    trie = OracleLookaheadCache(eos_ids=[2])
    ps = prompts(pseed, 3, 12, 64)
    batch = torch.cat(ps, 0)
    refs = [greedy_generate(model, p, max_new_tokens=30)['sequences'][0].tolist() for p in ps]
    assert all(0 not in r[12:] for r in refs)
    for rep in range(2):
        out = lookahead_generate_batch(model, trie, batch, max_new_tokens=30, decoding_length=48, branch_length=6)
        for b in range(3):
            got = out['sequences'][b, :out['lengths'][b]].tolist()
            assert got == refs[b], (family, rep, b, got, refs[b])
        assert sum(out['edls']) == sum(out['lengths']) - 3 * 12
        if rep == 1:  # the second pass drafts the first pass's answers from the per-slot output tries
            assert max(out['edls']) > 1, out['edls']
