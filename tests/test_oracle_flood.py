# -*- coding: utf-8 -*-
"""oracle/flood_draft.py (numpy restatement of FLOOD's hash-table lookahead draft, flood/flood/ops/draft.py) against
op streams recorded from the reference's own Triton kernels run under the Triton interpreter."""
import numpy as np

from oracle import flood_draft as F
from tests import flood_golden as G


class OracleImpl(object):
    def reset(self, T, BL, BC, V):
        self.T, self.BL, self.BC, self.V = T, BL, BC, V
        self.freq = np.zeros((T,), dtype=np.float32)
        self.table = np.zeros((T, BL), dtype=np.int32)

    def update(self, tokens):
        F.update_draft_table(tokens, self.freq, self.table, self.T, self.BL, self.BC, self.V)

    def retrieve(self, queries, RC):
        return F.retrieve_draft_table(queries, self.freq, self.table, self.T, self.V, self.BL, self.BC, RC)

    def verify(self, inp, nxt, offs, bs, RC):
        return F.verify_draft(inp, nxt, offs, bs, RC, self.BL)

    def cache_move(self, cache, src, dst):
        F.update_draft_cache(cache, src, dst)
        return cache

    def tables(self):
        return self.freq, self.table


def test_flood_oracle_reproduces_the_reference_kernels():
    n = 0
    for what, got, want in G.replay(OracleImpl()):
        assert got.shape == want.shape and np.array_equal(got.astype(want.dtype), want), what
        n += 1
    assert n > 150
