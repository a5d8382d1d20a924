# -*- coding: utf-8 -*-
"""CPU/torch restatement of the reference's BATCHED draft -> verify -> accept loop -- TEST INFRASTRUCTURE ONLY
(the next row to build; nothing on the product path imports this).

Follows /root/reference/lookahead/lookahead/common/pretrained_model_batch.py:

    draft (bat_get, sub-length = decoding_length // active requests)   :705-735
    accept walk per request (+ the max_length bound of the walk)       :810-905
    KV compaction per request                                          :907-918, :986-989
    loop: input puts ids[1:-1] per slot, cursors, stream_put per slot,
          stop rules, early stop of finished requests, final flush     :1193-1300

The reference runs one batched forward over a shared preallocated cache and places request b's draft rows at its own
cursor (models/llama/modeling_llama_batch.py:385-410).  Requests do not interact inside the forward (row b only sees
cache b), so the restatement drives one single-request backend per request (oracle.loop.HFBackend: installed Hugging
Face eager model, rank-4 0/1 mask -> additive mask, positions = row sum - 1): request b's mask row is
[ones(cursor_b) | its slice of the bat_get mask], which is exactly the columns the batched mask
cat([full[:, :, min_cur:min_cur+n, :min_cur], decoding_masks]) (:729-731) makes visible to it.

Pinned: tests/golden/batchloop_*.npz hold verify-step records of the reference's OWN batched loop code (imported from
/root/reference with four unused transformers names stubbed, tests/golden/ref_loop.py) and tests/test_loop_golden.py
replays them through this restatement: drafts, accepted tokens, dls, edls, the order requests leave the batch and
the padded output must be identical; plus the lossless property (tests/test_oracle_loop.py) and the bat_get golden
streams recorded from the live reference trie (tests/golden/trie_small_v30_batch.json)."""
import numpy as np
import torch

from .loop import HFBackend, _penalise


def _accept_batch(draft_ids, draft_masks, logits, seq, cur, max_length, penalty):
    """:810-905 for one request. draft_ids = decoding_ids[b][1:] (padding zeros included), draft_masks =
    decoding_masks[b, 1:, cur-min_cur+1:], logits [1, n, V], seq = LongTensor [1, max_length] (tokens beyond the
    cursor are pad), cur = index of the last known token.  Writes the accepted tokens into seq like the reference
    (:866).  returns (tokens, logit_indices)"""
    if len(draft_ids) == 0:  # :813-826
        scores = _penalise(logits[:, 0], seq[:, :cur + 1], penalty)
        nxt = int(torch.argmax(scores, dim=-1))
        seq[0, cur + 1] = nxt
        return [nxt], [0]
    branch_lengths = np.sum(draft_masks, axis=1)
    max_branch_length = int(np.max(branch_lengths))
    leaf_indices = [i - 1 for i, l in enumerate(branch_lengths) if not (i == 0 or l > branch_lengths[i - 1])] + [-1]
    leaf_masks = draft_masks[leaf_indices]
    leaf_lengths = branch_lengths[leaf_indices]
    seg = [0] + np.cumsum(leaf_lengths).tolist()
    nz = np.nonzero(leaf_masks)[1].tolist()
    mask_indices = [nz[seg[i]:seg[i + 1]] for i in range(len(seg) - 1)]
    branches = [[draft_ids[i] for i in idx] for idx in mask_indices]
    tokens, logit_indices = [], []
    for i in range(-1, min(max_branch_length, max_length - cur - 2)):  # :853: the walk never writes past max_length
        logit_index = 0 if i == -1 else mask_indices[0][i] + 1
        scores = _penalise(logits[:, logit_index], seq[:, :cur + i + 2], penalty)
        nxt = int(torch.argmax(scores, dim=-1))
        seq[0, cur + i + 2] = nxt
        tokens.append(nxt)
        logit_indices.append(logit_index)
        if i == max_branch_length - 1:
            break
        keep = [j for j, b in enumerate(branches) if len(b) > i + 1 and b[i + 1] == nxt]
        if not keep:
            break
        mask_indices = [mask_indices[j] for j in keep]
        branches = [branches[j] for j in keep]
    return tokens, logit_indices


@torch.no_grad()
def lookahead_generate_batch(model, trie, input_ids, max_new_tokens=None, max_length=None, eos_token_id=(2,),
                             decoding_length=64, branch_length=8, decoding_mode='hier', repetition_penalty=1.0,
                             pad_token_id=2, stop_words=None, backend_factory=None, trace=False):
    """restates lookahead_generation of the batch variant (:1002-1330) for equal-length, unpadded prompts.
    input_ids: LongTensor [bs, len].  returns dict(sequences [bs, <= max_length] padded with pad_token_id like the
    reference's output_ids, lengths = tokens known per request, dls, edls)"""
    bs, prompt_len = input_ids.shape
    dev = input_ids.device
    if max_length is None:
        max_length = prompt_len + int(max_new_tokens)
    eos = list(eos_token_id) if eos_token_id is not None else []
    trie.eos_ids = eos if eos else None                                       # :1141-1142
    trie.stop_words = stop_words if stop_words is not None else {}
    for i, ids in enumerate(input_ids.tolist()):                              # :1203-1206
        trie.put(ids[1:-1], branch_length=branch_length + 1, mode='input', idx=i)
    backends = [backend_factory(b) if backend_factory is not None else HFBackend(model) for b in range(bs)]
    steps = []
    seqs = [torch.cat([input_ids[b:b + 1], torch.full((1, max_length - prompt_len), pad_token_id, dtype=torch.long,
                                                      device=dev)], 1) for b in range(bs)]
    dls, edls = [], []
    # prefill (:781-808): mask = tril, first token = arg-max of the last prompt row, cursor = prompt length
    cursors = []
    for b in range(bs):
        m01 = torch.tril(torch.ones((1, 1, prompt_len, prompt_len), dtype=torch.long, device=dev))
        pos = (m01.sum(-1).squeeze(1) - 1).clamp(min=0)
        logits = backends[b].forward(input_ids[b:b + 1], m01, pos)
        scores = _penalise(logits[:, -1], input_ids[b:b + 1], repetition_penalty)
        seqs[b][0, prompt_len] = int(torch.argmax(scores, dim=-1))
        cursors.append(prompt_len)
        dls.append(1)
        edls.append(1)
    active = list(range(bs))                                                  # batch_indices (:1213)
    first = [[int(seqs[b][0, prompt_len])] for b in range(bs)]
    for b in range(bs):                                                       # :1243-1248 after the prefill step
        trie.stream_put(first[b], branch_length=branch_length + 1, final=False, mode='output', idx=b)
    finished = [b for b in active if cursors[b] + 1 >= max_length or any(e in first[b] for e in eos)]  # :1274-1281
    active = [b for b in active if b not in finished]
    mode = decoding_mode if '_' in decoding_mode else decoding_mode + '_mix'  # :709-711
    fmt, tmode = mode.split('_')
    while active:
        curs = [cursors[b] for b in active]
        qids = [[int(seqs[b][0, cursors[b] - 1]), int(seqs[b][0, cursors[b]])] for b in active]   # :705-707
        sub_dl = max(decoding_length // len(active), 1)                       # :712
        ids_list, masks, _sizes = trie.bat_get(qids, decoding_length=sub_dl, branch_length=branch_length,
                                               decoding_cursors=curs, mode=tmode, indices=list(active),
                                               decoding_mode=fmt)              # :714-720
        n = len(ids_list[0])
        assert all(len(x) == n for x in ids_list)                             # :722-723
        min_cur = min(curs)
        step_tokens = []
        for k, b in enumerate(active):
            cur = cursors[b]
            # request b's view of the batched mask (:729-731): columns [0, cur) = cached tokens, then its n draft
            # columns starting at its own cursor (bat_get places the tree at offset cur - min_cur)
            tree = torch.from_numpy(np.ascontiguousarray(masks[k][:, cur - min_cur:cur - min_cur + n])[None, None]).to(dev)
            m01 = torch.cat([torch.ones((1, 1, n, cur), dtype=torch.long, device=dev), tree.long()], dim=3)
            pos = (m01.sum(-1).squeeze(1) - 1).clamp(min=0)
            ids_in = torch.tensor([ids_list[k]], dtype=torch.long, device=dev)
            assert backends[b].rows() == cur
            logits = backends[b].forward(ids_in, m01, pos)
            draft_ids = ids_list[k][1:]
            draft_masks = masks[k][1:, cur - min_cur + 1:]
            tokens, logit_indices = _accept_batch(draft_ids, draft_masks, logits, seqs[b], cur, max_length,
                                                  repetition_penalty)
            dls.append(len(draft_ids) + 1)
            edls.append(len(tokens))
            # :907-918 keep rows [0, cur] + the accepted draft rows
            kv_idx = [li - 1 + cur + 1 for li in logit_indices[1:]]
            keep = list(range(cur + 1)) + kv_idx
            if len(keep) != backends[b].rows():
                backends[b].compact(torch.tensor(keep, dtype=torch.long, device=dev))
            cursors[b] = cur + len(tokens)
            step_tokens.append(tokens)
        if trace:
            steps.append(dict(active=list(active), ids=[list(x) for x in ids_list], cursors=curs, tokens=step_tokens))
        for k, b in enumerate(active):                                        # :1243-1248
            trie.stream_put(step_tokens[k], branch_length=branch_length + 1, final=False, mode='output', idx=b)
        still = []
        for k, b in enumerate(active):                                        # :1274-1283
            done = cursors[b] + 1 >= max_length or any(e in step_tokens[k] for e in eos)
            if not done:
                still.append(b)
        active = still
    for i in range(bs):                                                       # :1287-1289
        trie.stream_put([], branch_length=branch_length + 1, final=True, mode='output', idx=i)
    lengths = [cursors[b] + 1 for b in range(bs)]
    max_cur = max(cursors)
    res = dict(sequences=torch.cat([s[:, :max_cur + 1] for s in seqs], 0), lengths=lengths, dls=dls, edls=edls)
    if trace:
        res['steps'] = steps
    return res
