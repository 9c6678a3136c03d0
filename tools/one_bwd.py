#!/usr/bin/env python
"""A few block_grad launches at (M, N) for ncu launch lists / quick timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mamba_clip_b200 import _cabi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
M = int(sys.argv[2]) if len(sys.argv) > 2 else N
D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
be = _cabi.CudaBackend(path=2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.nn.functional.normalize(torch.randn(M, D, device="cuda", generator=g), dim=-1).bfloat16()
y = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=-1).bfloat16()
ls = torch.tensor([14.2857], device="cuda"); go = torch.ones(1, device="cuda")
lx = be.row_lse(x, y, ls, 0, False)[0]; ly = be.row_lse(y, x, ls, 0, False)[0]
for _ in range(3):
    be.block_grad(x, y, ls, go, lx, ly, 0, 1.0, 1.0, 2.0, 0.5 / N, False)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); be.block_grad(x, y, ls, go, lx, ly, 0, 1.0, 1.0, 2.0, 0.5 / N, False); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"block_grad M={M} N={N} D={D}: {sum(ts)/len(ts)*1e3:.1f} us (min {min(ts)*1e3:.1f})  persist={os.environ.get('MCLIP_BWD_PERSIST','1')}")
