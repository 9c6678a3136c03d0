#!/usr/bin/env python
"""GPU timeline of one ClipLoss fwd+bwd step (torch profiler / CUPTI), rank 0.  Launch with torchrun for W > 1:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_step.py [global_batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from mamba_clip_b200 import ClipLoss
from oracle import clip_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
Bl = B // world
img, txt = O.make_features(B, 512, seed=1234, dtype=torch.bfloat16)
a = img[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
b = txt[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
ls = torch.tensor(14.2857, device=dev, requires_grad=True)
crit = ClipLoss(True, True, True, rank, world)

def step():
    a.grad = b.grad = ls.grad = None
    crit(image_features=a, text_features=b, logit_scale=ls)["contrastive_loss"].backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    n = len(evs) // 3
    last_end = None
    print(f"world={world} B={B}: {len(evs)} GPU activities in 3 steps; middle step:")
    for e in evs[n:2 * n]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = (e.time_range.start - last_end) if last_end is not None else 0
        last_end = e.time_range.end
        print(f"  t={s:9.1f} us  dur={d:8.1f}  gap={gap:7.1f}  {e.name[:90]}")
    step_us = (evs[2 * n].time_range.start - evs[n].time_range.start)
    print(f"step period ~ {step_us:.1f} us")
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith(("ClipLossFunction", "aten::empty", "c10d", "nccl"))]
    tot = {}
    for e in cpu:
        tot[e.name] = tot.get(e.name, 0) + (e.time_range.end - e.time_range.start)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  cpu {k[:60]:60s} {v / 3:9.1f} us/step")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
