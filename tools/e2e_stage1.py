#!/usr/bin/env python
"""C5 of BASELINE.json: an end-to-end stage-1 training step with the fused loss dropped in, timed beside the same step
with the materialising reference formulation of the loss.

    python tools/e2e_stage1.py [--steps 10] [--batch 64] [--small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_stage1.py

The step mirrors reference src/mamba_clip/train.py:124-197,59-63,292-314 with `accum_freq == 1`:
    optimizer.zero_grad(); with autocast(bf16): model_out = model(images, texts); losses = loss(**model_out, target=targets);
    total_loss = sum(losses.values()); total_loss.backward(); optimizer.step(); logit_scale.clamp_(0, ln 100)
The towers are random-init stand-ins of the same architecture class as BiomedCLIP (ViT-B/16 at 224 px from `transformers`
+ a BERT-base text tower at 256 tokens, each with a linear projection to 512) -- there is no network for checkpoints, and the
reference's own model code (/root/reference) does not exist on the GPU box.  `ClipModel.forward`'s contract is kept
(reference model.py:1019-1064): F.normalize'd features + logit_scale.exp() in a dict that is splatted into the loss.
Loss A = mamba_clip_b200.ClipLoss (this repo), loss B = the materialising formulation (oracle/_ref ClipLoss when the copy
is present, else the same torch ops); both under DDP for W > 1 with (local_loss=False, gather_with_grad=False), the
configuration of the repo's SLURM script.  Prints one JSON line on rank 0: step times, the time of the loss fwd+bwd alone
(CUDA events around it in a separate pass with the towers' outputs detached and cached), and the loss's share of the step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class StandInClip(torch.nn.Module):
    """Same output contract as the reference's ClipModel.forward (model.py:1019-1064)."""

    def __init__(self, small: bool):
        super().__init__()
        from transformers import BertConfig, BertModel, ViTConfig, ViTModel
        if small:   # CPU-sized smoke configuration
            vc = ViTConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128, image_size=32, patch_size=16)
            tc = BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128, vocab_size=1000,
                            max_position_embeddings=64)
        else:       # ViT-B/16 @ 224 and BERT-base @ 256 tokens (PubMedBERT's shape)
            vc = ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224,
                           patch_size=16)
            tc = BertConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, vocab_size=30522,
                            max_position_embeddings=512)
        self.visual = ViTModel(vc, add_pooling_layer=False)
        self.text = BertModel(tc, add_pooling_layer=False)
        self.visual_proj = torch.nn.Linear(vc.hidden_size, 512, bias=False)
        self.text_proj = torch.nn.Linear(tc.hidden_size, 512, bias=False)
        self.logit_scale = torch.nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    def forward(self, image, text):
        vi = self.visual_proj(self.visual(pixel_values=image).last_hidden_state[:, 0])
        ti = self.text_proj(self.text(input_ids=text).last_hidden_state[:, 0])
        return {"image_features": F.normalize(vi, dim=-1), "text_features": F.normalize(ti, dim=-1),
                "logit_scale": self.logit_scale.exp()}


def reference_loss_module(rank, world):
    from oracle import build_ref
    if build_ref.available():
        return build_ref.load().ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank, world_size=world), "oracle/_ref"

    class Materialising(torch.nn.Module):       # same torch ops as reference loss.py:102-111,142-145 at W = 1
        def forward(self, image_features, text_features, logit_scale, target=None):
            lpi = logit_scale * image_features @ text_features.T
            lab = torch.arange(lpi.shape[0], device=lpi.device)
            return {"contrastive_loss": (F.cross_entropy(lpi, lab) + F.cross_entropy(lpi.T, lab)) / 2}
    if world > 1:
        raise SystemExit("the materialising loss needs oracle/_ref for world_size > 1")
    return Materialising(), "torch ops"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="per-rank batch (README: 64)")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="smoke run of the script's plumbing with the materialising loss only")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cpu") if args.cpu else torch.device("cuda", local_rank)
    if not args.cpu:
        torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo" if args.cpu else "nccl", rank=rank, world_size=world,
                                **({} if args.cpu else {"device_id": dev}))
    torch.manual_seed(1234)
    model = StandInClip(args.small).to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=None if args.cpu else [local_rank]) if world > 1 else model
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
    res = 32 if args.small else 224
    ntok = 64 if args.small else 256
    vocab = 1000 if args.small else 30522
    g = torch.Generator().manual_seed(100 + rank)
    images = torch.randn(args.batch, 3, res, res, generator=g).to(dev)
    texts = torch.randint(0, vocab, (args.batch, ntok), generator=g).to(dev)
    targets = torch.zeros(args.batch, dtype=torch.long, device=dev)
    losses = {}
    ref_mod, ref_kind = reference_loss_module(rank, world)
    losses["reference"] = ref_mod
    if not args.cpu:
        from mamba_clip_b200 import ClipLoss
        losses["fused"] = ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank, world_size=world)

    def sync():
        if not args.cpu:
            torch.cuda.synchronize(dev)

    def one_step(crit):
        opt.zero_grad()
        with torch.autocast(dev.type, dtype=torch.bfloat16):
            out = ddp(images, texts)
            ls = crit(**out, target=targets)
            total = sum(ls.values())
        total.backward()
        opt.step()
        with torch.no_grad():
            model.logit_scale.clamp_(0, math.log(100))
        return total

    import time
    result = {"config": {"workload": "C5 end-to-end stage-1 step, stand-in ViT-B/16 + BERT-base-256 towers, random init, synthetic 224px images / 256 tokens"
                                     if not args.small else "C5 smoke (tiny towers)",
                         "per_gpu_batch": args.batch, "world": world, "precision": "amp bf16", "reference_loss": ref_kind}}
    for name, crit in losses.items():
        for _ in range(args.warmup):
            one_step(crit)
        sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            total = one_step(crit)
        sync()
        dt = (time.perf_counter() - t0) / args.steps
        # the loss alone: fwd + bwd on detached, cached tower outputs (same dtype path as inside the step)
        with torch.no_grad(), torch.autocast(dev.type, dtype=torch.bfloat16):
            out = ddp(images, texts)
        feats = {k: v.detach().clone().requires_grad_(True) for k, v in out.items()}
        lt = []
        for it in range(args.warmup + args.steps):
            for v in feats.values():
                v.grad = None
            sync()
            if args.cpu:
                a = time.perf_counter()
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                lv = sum(crit(**feats, target=targets).values())
            lv.backward()
            if args.cpu:
                ms = (time.perf_counter() - a) * 1e3
            else:
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1)
            if it >= args.warmup:
                lt.append(ms)
        loss_ms = sorted(lt)[len(lt) // 2]
        result[name] = {"step_ms": dt * 1e3, "samples_per_s": args.batch * world / dt, "loss_fwd_bwd_ms": loss_ms,
                        "loss_share_of_step": loss_ms / (dt * 1e3), "last_loss": float(total.detach())}
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
