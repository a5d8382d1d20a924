# -*- coding: utf-8 -*-
"""Generates tests/golden/flood_draft.npz by running the REFERENCE's own Triton kernels
(/root/reference/flood/flood/ops/draft.py) under the Triton interpreter on the CPU.  Build container only:

    TRITON_INTERPRET=1 python tests/golden/gen_flood_golden.py

A seeded op stream over FLOOD's hash-table draft: update_draft_table (update_state), retrieve_draft_table
(proposal_draft), verify_draft, update_draft_cache, two table geometries (branch_count == branch_length and FLOOD's
default 8 / 32).  Every op is stored with what the reference returned, plus the final tables."""
import os
import sys

os.environ['TRITON_INTERPRET'] = '1'
import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.path.insert(0, '/root/reference/flood')
import flood.ops.draft as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def stream(tag, T, BL, BC, V, rng, n_ops, arrays, ops):
    freq = torch.zeros((T,), dtype=torch.float32)
    table = torch.zeros((T, BL), dtype=torch.int32)
    docs = []
    k = 0
    for _ in range(n_ops):
        kind = rng.choice(['update', 'update', 'retrieve', 'verify'])
        if kind == 'update' or not docs:
            if docs and rng.random() < 0.5:                       # repeat text: frequencies climb the 0.5..64 ladder
                toks = list(docs[int(rng.integers(0, len(docs)))])
                cut = int(rng.integers(0, max(len(toks) - 6, 1)))
                toks = toks[cut:]
            else:
                toks = rng.integers(3, 40, size=int(rng.integers(2, 60))).tolist()
            docs.append(toks)
            ref.update_draft_table(list(toks), freq, table, table_size=T, branch_length=BL, branch_count=BC, vocab=V)
            ops.append([tag, 'update', k])
            arrays[f'{tag}_{k}_tokens'] = np.array(toks, dtype=np.int32)
        elif kind == 'retrieve':
            RC = int(rng.choice([1, 2, 4, min(8, BC)]))
            qs = []
            for _ in range(int(rng.integers(1, 5))):
                d = docs[int(rng.integers(0, len(docs)))]
                j = int(rng.integers(0, max(len(d) - 1, 1)))
                qs.append(d[j:j + 2] if len(d) >= 2 and rng.random() < 0.85 else rng.integers(3, 40, size=2).tolist())
            out, masks = ref.retrieve_draft_table([list(q) for q in qs], freq, table, table_size=T, vocab=V,
                                                  branch_length=BL, branch_count=BC, retrieve_count=RC)
            ops.append([tag, 'retrieve', k, RC])
            arrays[f'{tag}_{k}_queries'] = np.array(qs, dtype=np.int32)
            arrays[f'{tag}_{k}_out'] = out.numpy().copy()
            arrays[f'{tag}_{k}_masks'] = masks.numpy().copy()
        else:
            RC, bs = int(rng.choice([2, 4])), int(rng.integers(1, 4))
            d = docs[int(rng.integers(0, len(docs)))]
            j = int(rng.integers(0, max(len(d) - 1, 1)))
            q = d[j:j + 2] if len(d) >= 2 else [3, 4]
            out, _ = ref.retrieve_draft_table([list(q)] * bs, freq, table, table_size=T, vocab=V, branch_length=BL,
                                              branch_count=BC, retrieve_count=RC)
            inp = out.reshape(-1).clone()
            # "model output": mostly the draft continued (token after position i = draft token i+1 of its branch)
            nxt = torch.zeros_like(inp)
            flat = out.numpy()
            for b in range(bs):
                row = flat[b]
                for br in range(RC):
                    for c in range(BL):
                        pos = br * BL + c
                        if pos >= RC * BL:
                            continue
                        follow = row[pos + 1] if pos + 1 < RC * BL else 0
                        nxt[b * RC * BL + pos] = int(follow) if rng.random() < 0.8 else int(rng.integers(3, 40))
                # position 0 (root) predicts the first token of SOME branch
                nxt[b * RC * BL] = int(row[1 + int(rng.integers(0, RC)) * BL]) if rng.random() < 0.9 else 1
            offs = torch.tensor([100 * b + 7 for b in range(bs)], dtype=torch.int32)
            o, s, dd = ref.verify_draft(inp, nxt, offs, None, bs, RC, BL)
            ops.append([tag, 'verify', k, RC, bs])
            arrays[f'{tag}_{k}_input'] = inp.numpy().copy()
            arrays[f'{tag}_{k}_next'] = nxt.numpy().copy()
            arrays[f'{tag}_{k}_offs'] = offs.numpy().copy()
            arrays[f'{tag}_{k}_vout'] = o.numpy().copy()
            arrays[f'{tag}_{k}_vsrc'] = s.numpy().copy()
            arrays[f'{tag}_{k}_vdst'] = dd.numpy().copy()
            cache = torch.arange(400 * 6, dtype=torch.float32).view(400, 6).clone()
            ref.update_draft_cache(cache, s, dd)
            arrays[f'{tag}_{k}_cache'] = cache.numpy().copy()
        k += 1
    arrays[f'{tag}_final_freq'] = freq.numpy().copy()
    arrays[f'{tag}_final_table'] = table.numpy().copy()
    arrays[f'{tag}_geom'] = np.array([T, BL, BC, V], dtype=np.int64)


def main():
    rng = np.random.default_rng(5)
    arrays, ops = {}, []
    stream('sq', 1 << 10, 8, 8, 50, rng, 70, arrays, ops)         # branch_count == branch_length, crowded table
    stream('wide', 1 << 12, 8, 32, 4000, rng, 60, arrays, ops)    # FLOOD's default 8 / 32 geometry
    import json
    arrays['ops'] = np.frombuffer(json.dumps(ops).encode(), dtype=np.uint8)
    path = os.path.join(HERE, 'flood_draft.npz')
    np.savez_compressed(path, **arrays)
    print(len(ops), 'ops', os.path.getsize(path) // 1024, 'KB',
          'nonzero freq', int((arrays['sq_final_freq'] > 0).sum()), int((arrays['wide_final_freq'] > 0).sum()),
          'max freq', float(arrays['sq_final_freq'].max()))


if __name__ == '__main__':
    main()
