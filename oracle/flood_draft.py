# -*- coding: utf-8 -*-
"""numpy restatement of FLOOD's hash-table lookahead draft -- TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu leg).

Follows /root/reference/flood/flood/ops/draft.py: update_draft_table :168-204 (kernel :92-165), retrieve_draft_table
:352-402 (kernel :278-349), verify_draft :491-543 (kernel :406-488), update_draft_cache :562-570 (kernel :547-559).
Pinned: tests/golden/flood_draft.npz holds op streams recorded by running the reference's own Triton kernels under
TRITON_INTERPRET=1 (tests/golden/gen_flood_golden.py); tests/test_oracle_flood.py replays them bit for bit.
The reference's update kernel races between its programs on a GPU; the interpreter runs them in order, which is the
sequential semantics restated here (and implemented by csrc/flood_draft.cu)."""
import numpy as np


def update_draft_table(tokens, freq_table, draft_table, table_size, branch_length, branch_count, vocab):
    """in place on freq_table float32 [size], draft_table int32 [size, BL]"""
    n = len(tokens)
    if n <= 3:
        return
    BL, BC = branch_length, branch_count
    for p in range(0, n - 3):                                   # p + 4 <= token_count (:108)
        bucket = (int(tokens[p]) * vocab + int(tokens[p + 1])) % (table_size - BC)
        branch = np.array([tokens[p + 2 + d] if p + 2 + d < n else 0 for d in range(BL)], dtype=np.int32)
        uid = np.int32(branch.sum(dtype=np.int32))
        hit = False
        for j in range(BC):                                     # :127-148
            freq = freq_table[bucket + j]
            match = (not hit) and np.int32(draft_table[bucket + j].sum(dtype=np.int32)) == uid
            empty = (not hit) and freq == 0
            freq_table[bucket + j] = freq + 1.0 if match else (1.0 if empty else freq)
            if empty:
                draft_table[bucket + j] = branch
            hit = hit or match or empty
        not_hit = not hit
        for j in range(BC):                                     # :152-165
            half = np.float32(freq_table[bucket + j]) / np.float32(2.0)
            replace = half < 1.0 and not_hit
            freq_table[bucket + j] = 1.0 if replace else half
            if replace:
                draft_table[bucket + j] = branch


def retrieve_draft_table(queries, freq_table, draft_table, table_size, vocab, branch_length, branch_count,
                         retrieve_count):
    """queries: [[p0, p1], ...] -> (tokens int32 [bs, RC*BL], masks int8 [bs, RC*BL, RC*BL])"""
    BL, BC, RC = branch_length, branch_count, retrieve_count
    bs = len(queries)
    ln = RC * BL
    out = np.zeros((bs, ln + 1), dtype=np.int32)
    masks = np.tril(np.ones((bs, ln, ln), dtype=np.int8))
    for j in range(1, RC):
        masks[:, j * BL + 1:(j + 1) * BL + 1, 1:j * BL + 1] = 0
    for b, (p0, p1) in enumerate(queries):
        bucket = (int(p0) * vocab + int(p1)) % (table_size - BC)
        freqs = freq_table[bucket:bucket + BL]                 # `indices = arange(BRANCH_LENGTH)` (:290)
        done = False
        for i in range(8):
            valid = freqs >= 2.0 ** (8 - i - 2)
            cs = np.cumsum(valid)
            sel = valid & (cs <= RC)
            if sel.sum() >= RC:
                for j in np.flatnonzero(sel):
                    out[b, 1 + (cs[j] - 1) * BL:1 + cs[j] * BL] = draft_table[bucket + j]
                done = True
                break
        if not done:
            valid = freqs >= 0.5
            cs = np.cumsum(valid)
            sel = valid & (cs <= RC)
            for j in np.flatnonzero(sel):
                out[b, 1 + (cs[j] - 1) * BL:1 + cs[j] * BL] = draft_table[bucket + j]
        out[b, 0] = p1
    return out[:, :-1].copy(), masks


def verify_draft(input_ids, next_ids, cache_offsets, batch_size, branch_count, branch_length):
    BC, BL = branch_count, branch_length
    inp = np.asarray(input_ids, dtype=np.int64).reshape(batch_size, BC, BL)
    nxt = np.asarray(next_ids, dtype=np.int64).reshape(batch_size, BC, BL)

    def tiles(a):
        t = -np.ones((batch_size, BC, BL + 1), dtype=np.int64)
        t[:, :, 0] = a[:, 0:1, 0]
        t[:, :, 1:-1] = a[:, :, 1:]
        t[:, :-1, -1] = a[:, 1:, 0]
        return t

    ti, tn = tiles(inp), tiles(nxt)
    out = -np.ones((batch_size, BL + 1), dtype=np.int64)
    src = -np.ones((batch_size * BL,), dtype=np.int64)
    dst = -np.ones((batch_size * BL,), dtype=np.int64)
    for b in range(batch_size):
        best, best_i = 0, 0
        for i in range(BC):
            acc = 0
            for j in range(BL):
                if ti[b, i, j + 1] == tn[b, i, j]:
                    acc += 1
                    if acc > best:
                        best, best_i = acc, i
                else:
                    break
        for j in range(BL):
            if j == 0:
                out[b, 0] = tn[b, best_i, 0]
            if ti[b, best_i, j + 1] != tn[b, best_i, j]:
                break
            out[b, j + 1] = tn[b, best_i, j + 1]
            src[b * BL + j] = cache_offsets[b] + BL * best_i + 1 + j
            dst[b * BL + j] = cache_offsets[b] + 1 + j
    return out, src, dst


def update_draft_cache(cache, src, dst):
    for s, d in zip(src, dst):
        if s >= 0 and d != s:
            cache[d] = cache[s]
