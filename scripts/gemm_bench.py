# -*- coding: utf-8 -*-
"""pia weight-streaming GEMM vs cuBLAS (torch.mm) at the Llama-2-7B decode shapes, 32 distinct weights per shape
(so no weight is L2 resident), CUDA-graph replay, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from painlessinferenceacceleration_b200.common import ops  # noqa: E402

dev = 'cuda:0'
NL = 32


def timeit(fn, per, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * per)


for name, N, K, splits in [('qkv', 12288, 4096, (1, 2)), ('o', 4096, 4096, (1, 2, 4, 8)), ('gate_up', 22016, 4096, (1, 2)),
                           ('down', 4096, 11008, (2, 4, 8)), ('lm_head', 32000, 4096, (1,))]:
    nl = NL if N * K < 1.5e8 else 8
    ws = [(torch.randn((N, K), device=dev) * 0.02).to(torch.bfloat16) for _ in range(nl)]
    x = torch.randn((64, K), device=dev).to(torch.bfloat16)
    by = N * K * 2
    t = timeit(lambda: [torch.mm(x, w.t()) for w in ws], nl)
    print(f'{name:8s} cuBLAS           {t:8.2f} us  {by / t / 1e3:7.0f} GB/s', flush=True)
    for tiled in (False, True):
        wt = [ops.tile_weight(w) for w in ws] if tiled else ws
        for s in splits:
            gs = [ops.Gemm(w, x, split_k=s, tiled=tiled) for w in wt]
            t = timeit(lambda: [g.run(64) for g in gs], nl)
            print(f'{name:8s} pia split_k={s:<2d} {"tiled" if tiled else "rowmj"} {t:8.2f} us  {by / t / 1e3:7.0f} GB/s', flush=True)
            del gs
        if tiled:
            gs = [ops.Gemm(w, x, split_k=-1, tiled=True) for w in wt]
            t = timeit(lambda: [g.run(64) for g in gs], nl)
            print(f'{name:8s} pia stream-K   tiled {t:8.2f} us  {by / t / 1e3:7.0f} GB/s', flush=True)
            del gs
        del wt
    del ws

# fused SiLU*up epilogue vs GEMM + k_silu_mul
N, K = 22016, 4096
ws = [(torch.randn((N, K), device=dev) * 0.02).to(torch.bfloat16) for _ in range(8)]
x = torch.randn((64, K), device=dev).to(torch.bfloat16)
act = torch.zeros((64, N // 2), dtype=torch.bfloat16, device=dev)
gu = torch.zeros((64, N), dtype=torch.bfloat16, device=dev)
g1 = [ops.Gemm(ops.tile_weight(w), x, tiled=True) for w in ws]
t = timeit(lambda: [(g.run(64, out=gu), ops.silu_mul(gu, act)) for g in g1], 8)
print(f'gate_up  pia + k_silu_mul      {t:8.2f} us', flush=True)
t = timeit(lambda: [(torch.mm(x, w.t(), out=gu), ops.silu_mul(gu, act)) for w in ws], 8)
print(f'gate_up  cuBLAS + k_silu_mul   {t:8.2f} us', flush=True)
g2 = [ops.Gemm(ops.tile_weight(ops.interleave_gate_up(w)), x, tiled=True).set_silu() for w in ws]
t = timeit(lambda: [g.run(64, out=act) for g in g2], 8)
print(f'gate_up  pia fused silu        {t:8.2f} us', flush=True)
del g1, g2, ws
# down: GEMM (+ split-K partials) followed by the rmsnorm that consumes it
N, K = 4096, 11008
ws = [(torch.randn((N, K), device=dev) * 0.02).to(torch.bfloat16) for _ in range(16)]
x = torch.randn((64, K), device=dev).to(torch.bfloat16)
r = torch.zeros((64, N), dtype=torch.bfloat16, device=dev)
y = torch.zeros_like(r)
wn = torch.ones((N,), dtype=torch.bfloat16, device=dev)
t = timeit(lambda: [(ops.rmsnorm(torch.mm(x, w.t()), r, wn, 1e-5, r, y)) for w in ws], 16)
print(f'down     cuBLAS + rmsnorm            {t:8.2f} us', flush=True)
for sk in (2, 4, 8):
    gs = [ops.Gemm(ops.tile_weight(w), x, split_k=sk, tiled=True) for w in ws]
    t = timeit(lambda: [ops.rmsnorm_partials(g.run(64), r, wn, 1e-5, r, y) for g in gs], 16)
    print(f'down     pia split_k={sk} + rmsnorm_partials {t:8.2f} us', flush=True)
    del gs
