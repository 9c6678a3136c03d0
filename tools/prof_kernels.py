#!/usr/bin/env python
"""Short program for ncu: two fwd+bwd steps of the headline config (B=32768, D=512, bf16, W=1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mamba_clip_b200 import ClipLoss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator(device="cuda").manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda", generator=g), dim=-1).bfloat16().requires_grad_(True)
txt = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda", generator=g), dim=-1).bfloat16().requires_grad_(True)
ls = torch.tensor(14.2857, device="cuda", requires_grad=True)
crit = ClipLoss()
for _ in range(2):
    img.grad = txt.grad = ls.grad = None
    loss = crit(img, txt, ls)["contrastive_loss"]
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()), "dls", float(ls.grad))
