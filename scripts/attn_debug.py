# -*- coding: utf-8 -*-
"""Phase timeline of k_tree_attn (globaltimer stamps written by the kernel when a debug buffer is attached)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from painlessinferenceacceleration_b200 import _lib as L  # noqa: E402
from painlessinferenceacceleration_b200.common import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--P', type=int, default=384)
ap.add_argument('--n', type=int, default=64)
ap.add_argument('--hq', type=int, default=32)
ap.add_argument('--hkv', type=int, default=32)
ap.add_argument('--max-seq', type=int, default=577)
ap.add_argument('--layers', type=int, default=32)
a = ap.parse_args()
dev = 'cuda:0'
D, R = 128, 64
kc = (torch.randn((a.layers, a.hkv, a.max_seq, D), device=dev) * 0.5).to(torch.bfloat16)
vc = (torch.randn((a.layers, a.hkv, a.max_seq, D), device=dev) * 0.5).to(torch.bfloat16)
q = (torch.randn((R, a.hq, D), device=dev) * 0.5).to(torch.bfloat16)
out = torch.zeros_like(q)
plan = ops.AttnPlan(kc, vc, a.hq, a.hkv, D, R)
ns, ng = C.c_int(0), C.c_int(0)
L.check(plan.lib.pia_attn_plan_grid(plan.h, C.byref(ns), C.byref(ng)))
print('grid', ns.value, 'x', ng.value)
mask = torch.zeros((R, 1), dtype=torch.int64, device=dev)
rows = np.array([(1 << (i + 1)) - 1 for i in range(R)], dtype=np.uint64)
mask[:, 0] = torch.from_numpy(rows.view(np.int64)).to(dev)
dn = torch.tensor([a.n], dtype=torch.int32, device=dev)
dP = torch.tensor([a.P], dtype=torch.int32, device=dev)
slots = ops.Slots(dn, dP, None, 64)
for li in range(a.layers):
    plan.forward(li, q, mask, slots, out)
torch.cuda.synchronize()
dbg = torch.zeros((ns.value * ng.value, 16), dtype=torch.int64, device=dev)
L.check(plan.lib.pia_attn_plan_set_debug(plan.h, dbg.data_ptr()))
names = ['start', 'setup done', 'tma0 issued', 'qk0 issued', 'mma done', 'q loaded', 's0 ready', 'p0 written',
         'o0 ready', 'tiles done', 'row written', 'cta end', 'combined(t0)', 'synced(t0)', 'barrier A', 'pushed']
for trial in range(3):
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.forward(trial + 3, q, mask, slots, out)
    e1.record()
    torch.cuda.synchronize()
    t = dbg.cpu().numpy().astype(np.int64)
    live = t[:, 0] > 0
    t = t[live]
    t0 = t[:, 0].min()
    print(f'trial {trial}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, {live.sum()} CTAs ran')
    for k, nm in enumerate(names):
        col = t[:, k]
        col = col[col > 0] - t0
        if len(col):
            print(f'  {nm:12s} min {col.min() / 1e3:7.2f}  med {np.median(col) / 1e3:7.2f}  max {col.max() / 1e3:7.2f} us')
L.check(plan.lib.pia_attn_plan_set_debug(plan.h, None))
