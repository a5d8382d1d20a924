# -*- coding: utf-8 -*-
"""k_gemm_ws alone on the Llama-2-7B gate/up projection (the step's dominant kernel), 8 distinct tiled weights so that
no launch is L2 resident, for `ncu --set full -k regex:k_gemm_ws`.  `silu` as argument: the fused SiLU*up epilogue;
`down`: the down projection (N=4096, K=11008) as the 4-CTA cluster split-K the decode step uses.
Numbers printed here are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from painlessinferenceacceleration_b200.common import ops  # noqa: E402

dev = 'cuda:0'
silu = len(sys.argv) > 1 and sys.argv[1] == 'silu'
down = len(sys.argv) > 1 and sys.argv[1] == 'down'
N, K = (4096, 11008) if down else (22016, 4096)
x = torch.randn((64, K), device=dev).to(torch.bfloat16)
ws = [(torch.randn((N, K), device=dev) * 0.02).to(torch.bfloat16) for _ in range(8)]
plans = [ops.Gemm(ops.tile_weight(ops.interleave_gate_up(w) if silu else w), x, tiled=True, split_k=-4 if down else 1)
         for w in ws]
out = torch.empty((64, N // 2 if silu else N), dtype=torch.bfloat16, device=dev)
if silu:
    [p.set_silu() for p in plans]
for rep in range(3):
    for p in plans:
        p.run(64, out=out)
torch.cuda.synchronize()
print('done', silu)
