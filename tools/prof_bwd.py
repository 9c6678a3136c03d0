#!/usr/bin/env python
"""One block_grad launch at a given size with MCLIP_DBG=16 (profile build): prints the in-kernel wait accounting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mamba_clip_b200 import _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
M = int(sys.argv[2]) if len(sys.argv) > 2 else B
be = _cabi.CudaBackend(path=2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.nn.functional.normalize(torch.randn(M, 512, device="cuda", generator=g), dim=-1).bfloat16()
y = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda", generator=g), dim=-1).bfloat16()
ls = torch.tensor([14.2857], device="cuda"); go = torch.ones(1, device="cuda")
lx = be.row_lse(x, y, ls, 0, False)[0]; ly = be.row_lse(y, x, ls, 0, False)[0]
for _ in range(2):
    be.block_grad(x, y, ls, go, lx, ly, 0, 1.0, 1.0, 2.0, 0.5 / B, False)
torch.cuda.synchronize()
