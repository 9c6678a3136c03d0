#!/usr/bin/env python
"""How long does the host need to enqueue one fwd+bwd step?  Small batch, so the GPU is never the limiter."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mamba_clip_b200 import ClipLoss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=-1).bfloat16().requires_grad_(True)
txt = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=-1).bfloat16().requires_grad_(True)
ls = torch.tensor(14.2857, device="cuda", requires_grad=True)
crit = ClipLoss()


def step():
    img.grad = txt.grad = ls.grad = None
    crit(img, txt, ls)["contrastive_loss"].backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, with final sync {1e3 * (t2 - t0) / n:.3f} ms/step")
if len(sys.argv) > 2:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
