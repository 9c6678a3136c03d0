#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (`/root/reference/src/mamba_clip/loss.py`).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes `tests/golden/single.npz` (W=1 cases) and `tests/golden/ranks.npz` (2/4-rank gloo runs of
the real `ClipLoss` for all four (local_loss, gather_with_grad) modes).  Inputs are NOT stored:
they regenerate from the seeds via `oracle.clip_oracle.make_features`.  Large gradients are stored
as a projection onto a fixed seeded [D, 8] matrix plus their Frobenius norm; small ones in full.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.distributed.nn  # noqa: F401  (loss.py:26 relies on this side-effect import)
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle.clip_oracle import make_features  # noqa: E402

FULL_LIMIT = 16384  # store full gradients when B*D <= this


def projector(dim: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(4242)
    return torch.randn(dim, 8, generator=g, dtype=torch.float64)


SINGLE_CASES = []
for B, D in [(1, 64), (2, 768), (7, 512), (32, 512), (64, 64), (64, 512), (100, 512), (129, 768), (512, 512)]:
    for ls in (1.0, 14.2857, 100.0):
        for corr in (False, True):
            if B >= 100 and ls == 1.0:
                continue
            SINGLE_CASES.append(dict(B=B, D=D, ls=ls, corr=corr, go=1.0 if not corr else 3.0,
                                     bf16=False, seed=1000 + B + D))
# bf16-valued inputs, reference run on their fp32 upcast
for B, D in [(64, 512), (129, 768), (512, 512)]:
    for ls in (14.2857, 100.0):
        SINGLE_CASES.append(dict(B=B, D=D, ls=ls, corr=True, go=2.0, bf16=True, seed=2000 + B))

RANK_CASES = []
for W, Bl, D in [(2, 16, 64), (4, 8, 32)]:
    for local_loss in (False, True):
        for gwg in (False, True):
            for corr in (False, True):
                RANK_CASES.append(dict(W=W, Bl=Bl, D=D, local_loss=local_loss, gwg=gwg, corr=corr,
                                       ls=30.0 if corr else 14.2857, go=3.0, seed=3000 + W * 10 + Bl))


def run_single(case):
    from mamba_clip.loss import ClipLoss
    img, txt = make_features(case["B"], case["D"], seed=case["seed"], correlated=case["corr"])
    if case["bf16"]:
        img = img.bfloat16().float()
        txt = txt.bfloat16().float()
    img.requires_grad_(True)
    txt.requires_grad_(True)
    ls = torch.tensor(case["ls"], dtype=torch.float32, requires_grad=True)
    loss = ClipLoss()(img, txt, ls)["contrastive_loss"]
    loss.backward(torch.tensor(case["go"]))
    return loss.detach(), img.grad, txt.grad, ls.grad


def pack_grad(out, key, g):
    g = g.double()
    out[key + "_norm"] = np.float64(g.norm())
    if g.numel() <= FULL_LIMIT:
        out[key + "_full"] = g.float().numpy()
    else:
        out[key + "_proj"] = (g @ projector(g.shape[1])).numpy()


def _rank_worker(rank, case, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=case["W"])
    from mamba_clip.loss import ClipLoss
    W, Bl = case["W"], case["Bl"]
    img, txt = make_features(W * Bl, case["D"], seed=case["seed"], correlated=case["corr"])
    img = img[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
    txt = txt[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
    ls = torch.tensor(case["ls"], dtype=torch.float32, requires_grad=True)
    crit = ClipLoss(local_loss=case["local_loss"], gather_with_grad=case["gwg"], cache_labels=True,
                    rank=rank, world_size=W)
    loss = crit(img, txt, ls)["contrastive_loss"]
    loss.backward(torch.tensor(case["go"]))
    q.put((rank, loss.item(), img.grad.numpy(), txt.grad.numpy(), float(ls.grad)))
    dist.barrier()
    dist.destroy_process_group()


def run_ranks(case, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, case, port, q)) for r in range(case["W"])]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join()
    return sorted(res, key=lambda t: t[0])


def main():
    torch.manual_seed(0)
    single = {"cases": json.dumps(SINGLE_CASES)}
    for k, case in enumerate(SINGLE_CASES):
        loss, di, dt, dls = run_single(case)
        single[f"c{k}_loss"] = np.float64(loss)
        single[f"c{k}_dls"] = np.float64(dls)
        pack_grad(single, f"c{k}_di", di)
        pack_grad(single, f"c{k}_dt", dt)
    np.savez_compressed(os.path.join(HERE, "single.npz"), **single)

    ranks = {"cases": json.dumps(RANK_CASES)}
    for k, case in enumerate(RANK_CASES):
        res = run_ranks(case, 29600 + k)
        for rank, loss, di, dt, dls in res:
            ranks[f"c{k}_r{rank}_loss"] = np.float64(loss)
            ranks[f"c{k}_r{rank}_dls"] = np.float64(dls)
            ranks[f"c{k}_r{rank}_di"] = di
            ranks[f"c{k}_r{rank}_dt"] = dt
    np.savez_compressed(os.path.join(HERE, "ranks.npz"), **ranks)
    print("single cases:", len(SINGLE_CASES), "rank cases:", len(RANK_CASES))
    for f in ("single.npz", "ranks.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
