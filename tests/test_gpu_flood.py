# -*- coding: utf-8 -*-
"""FLOOD `Spec` integration on the GPU: painlessinferenceacceleration_b200.flood.speculative (the
reference's flood/flood/utils/speculative.py surface over csrc/flood_draft.cu, through the C ABI) replays op streams
recorded from the reference's own Triton kernels bit for bit, and the `Lookahead(Spec)` class round-trips a
propose -> verify -> update_cache step like FLOOD's batcher drives it (flood/utils/batch.py:484)."""
import types

import numpy as np
import pytest
import torch

from tests import flood_golden as G

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class GpuImpl(object):
    def reset(self, T, BL, BC, V):
        self.T, self.BL, self.BC, self.V = T, BL, BC, V
        self.freq = torch.zeros((T,), dtype=torch.float32, device=DEV)
        self.table = torch.zeros((T, BL), dtype=torch.int32, device=DEV)

    def update(self, tokens):
        from painlessinferenceacceleration_b200.flood import speculative as S
        S.update_draft_table(list(tokens), self.freq, self.table, table_size=self.T, branch_length=self.BL,
                             branch_count=self.BC, vocab=self.V)

    def retrieve(self, queries, RC):
        from painlessinferenceacceleration_b200.flood import speculative as S
        out, masks = S.retrieve_draft_table([list(q) for q in queries], self.freq, self.table, table_size=self.T,
                                            vocab=self.V, branch_length=self.BL, branch_count=self.BC, retrieve_count=RC)
        return out.cpu().numpy(), masks.cpu().numpy()

    def verify(self, inp, nxt, offs, bs, RC):
        from painlessinferenceacceleration_b200.flood import speculative as S
        o, s, d = S.verify_draft(torch.from_numpy(inp).to(DEV), torch.from_numpy(nxt).to(DEV),
                                 torch.from_numpy(offs).to(DEV), None, bs, RC, self.BL)
        return o.cpu().numpy(), s.cpu().numpy(), d.cpu().numpy()

    def cache_move(self, cache, src, dst):
        from painlessinferenceacceleration_b200.flood import speculative as S
        c = torch.from_numpy(cache).to(DEV)
        S.update_draft_cache(c, torch.from_numpy(src).to(DEV), torch.from_numpy(dst).to(DEV))
        return c.cpu().numpy()

    def tables(self):
        return self.freq.cpu().numpy(), self.table.cpu().numpy()


def test_flood_kernels_reproduce_the_reference_kernels():
    n = 0
    for what, got, want in G.replay(GpuImpl()):
        assert got.shape == want.shape and np.array_equal(got.astype(want.dtype), want), what
        n += 1
    assert n > 150


def test_lookahead_spec_surface():
    """Lookahead(Spec): update_state -> proposal_draft -> verify_draft -> update_cache with FLOOD's argument shapes"""
    from painlessinferenceacceleration_b200.flood import Lookahead, Spec
    la = Lookahead(table_size=1 << 14, branch_length=8, branch_count=32, vocab_size=5000, device=torch.device(DEV))
    assert isinstance(la, Spec)
    text = [11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23]
    for _ in range(3):
        la.update_state(list(text))
    toks, masks = la.proposal_draft([[12, 13], [400, 401]], retrieve_count=4)
    assert toks.shape == (2, 32) and masks.shape == (2, 32, 32) and masks.dtype == torch.int8
    assert toks[0, :9].tolist() == [13, 14, 15, 16, 17, 18, 19, 20, 21]     # the continuation of "12 13"
    assert toks[1].tolist() == [401] + [0] * 31                              # unknown context: no branch
    # the model "continues the text": position i predicts draft token i+1
    nxt = torch.zeros_like(toks)
    nxt[:, :-1] = toks[:, 1:]
    meta = types.SimpleNamespace(batch_size=2, retrieve_count=4,
                                 cache_indices=torch.tensor([[50 + i for i in range(32)], [200 + i for i in range(32)]],
                                                            device=DEV, dtype=torch.int32))
    out, src, dst = la.verify_draft(toks.reshape(-1), nxt.reshape(-1), batch_meta_info=meta)
    assert out[0].tolist() == [14, 15, 16, 17, 18, 19, 20, 21, 22 if False else int(nxt[0, 8])]
    assert src[:8].tolist() == [51 + i for i in range(8)] and dst[:8].tolist() == [51 + i for i in range(8)]
    caches = types.SimpleNamespace(caches=[torch.arange(300 * 4, dtype=torch.bfloat16, device=DEV).view(300, 4).clone()],
                                   num_layers=1, fix_size_indices=None)
    before = caches.caches[0].clone()
    la.update_cache(src, dst, caches)
    assert torch.equal(caches.caches[0], before)      # branch 0 accepted: rows already in place
