"""A few mclip_fused_grad calls at (M, N) = (B, B), D = 512 bf16 -- for ncu launch lists / A-B timing of the
shared-recompute backward (tools/README.md).  usage: python tools/one_fused.py [B] [D] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_clip_b200 import _cabi  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402  (input generator only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
be = _cabi.get_backend()
img, txt = O.make_features(B, D, seed=1234, dtype=torch.bfloat16)
x, y = img.cuda(), txt.cuda()
ls = torch.full((1,), 14.2857, device="cuda")
go = torch.ones(1, device="cuda")
row_lse, _ = be.row_lse(x, y, ls, 0, False)
col_lse, _ = be.row_lse(y, x, ls, 0, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("fused_grad", lambda: be.fused_grad(x, y, ls, go, row_lse, col_lse, 0, 0.5 / B)),
                 ("block_grad x2", lambda: (be.block_grad(x, y, ls, go, row_lse, col_lse, 0, 1.0, 1.0, 2.0, 0.5 / B, False),
                                            be.block_grad(y, x, ls, go, col_lse, row_lse, 0, 1.0, 1.0, 2.0, 0.5 / B, True)))):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name}: B={B} D={D}  ms per call: " + " ".join(f"{t:.3f}" for t in ts), flush=True)
