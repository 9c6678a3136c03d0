"""Multi-GPU (NCCL) parity of ClipLoss: one process per GPU, all four (local_loss, gather_with_grad) modes,
against the golden vectors of the real reference run under gloo and, at tensor-core sizes, against the
single-process rank emulation of the oracle.  Skipped when fewer than 2 GPUs are visible."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _features(W, job):
    dtype = getattr(torch, job["dtype"])
    img, txt = O.make_features(W * job["Bl"], job["D"], seed=job["seed"], correlated=job["corr"])
    if job.get("adv"):
        # identical pairs (cos = 1), then the first half of the image rows shrunk: at ls = 100 every logit of the first
        # half's columns lies ~80 log2 units below the other half's positives, so the two-sided forward of the ranks
        # owning those columns leaves its f32 window and they must take the predicated one-sided path
        txt = img.clone()
        img[: W * job["Bl"] // 2] *= 0.01
    return img.to(dtype), txt.to(dtype)


def _worker(rank, W, port, jobs, q):
    import threading
    wd = threading.Timer(420.0, lambda: os._exit(3))        # a hung collective must not outlive the test
    wd.daemon = True
    wd.start()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=W, device_id=dev)
    try:
        import mamba_clip_b200
        from mamba_clip_b200 import ClipLoss
        for j, job in enumerate(jobs):
            # "graphs": the step is repeated so that the last repetition replays the captured CUDA graphs (collectives
            # included); every repetition must give the same result
            mamba_clip_b200.enable_cuda_graphs(bool(job.get("graphs")))
            # "no_small": keep the job on the general (tensor-core) path although its size qualifies for the latency path
            if job.get("no_small"):
                os.environ["MCLIP_NO_SMALL_PATH"] = "1"
            else:
                os.environ.pop("MCLIP_NO_SMALL_PATH", None)
            # "precopy": force the f16 operand copies onto the side stream next to the forward (default only from 8 ranks)
            mamba_clip_b200._function._F16_PRECOPY = True if job.get("precopy") else None
            reps = 5 if job.get("graphs") else 1
            Bl, D = job["Bl"], job["D"]
            dtype = getattr(torch, job["dtype"])
            img, txt = _features(W, job)
            a = img[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
            b = txt[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
            ls = torch.tensor(job["ls"], device=dev, requires_grad=True)
            crit = ClipLoss(job["local_loss"], job["gwg"], True, rank, W)
            if job.get("proj"):
                # raw projections -> fused normalise + cast into the gather slot -> in-place all-gather -> loss
                from mamba_clip_b200.producer import clip_loss_from_projections
                a = (a.detach().float() * 3.0).requires_grad_(True)
                b = (b.detach().float() * 0.5).requires_grad_(True)
            for _ in range(reps):
                a.grad = b.grad = ls.grad = None
                if job.get("proj"):
                    loss = clip_loss_from_projections(crit, a, b, ls, dtype=torch.bfloat16)["contrastive_loss"]
                else:
                    loss = crit(a, b, ls)["contrastive_loss"]
                loss.backward(torch.tensor(job["go"], device=dev))
            q.put((j, rank, float(loss.detach()), a.grad.float().cpu().numpy(), b.grad.float().cpu().numpy(), float(ls.grad)))
        mamba_clip_b200.enable_cuda_graphs(False)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(W, jobs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, W, port, jobs, q)) for r in range(W)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(W * len(jobs))]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return {(j, r): rest for j, r, *rest in res}


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_two_ranks_against_reference_golden():
    Z = np.load(os.path.join(GOLD, "ranks.npz"))
    cases = json.loads(str(Z["cases"]))
    ks = [k for k, c in enumerate(cases) if c["W"] == 2]
    jobs = [dict(cases[k], dtype="float32") for k in ks]
    out = _run(2, jobs)
    for j, k in enumerate(ks):
        c = cases[k]
        floor = 4 * 1.2e-7 * max(1.0, c["ls"]) * c["go"] * c["ls"] / (2 * c["Bl"]) * c["Bl"] ** 0.5
        for r in range(2):
            loss, di, dt, dls = out[(j, r)]
            gl, gd = float(Z[f"c{k}_r{r}_loss"]), float(Z[f"c{k}_r{r}_dls"])
            gi = torch.from_numpy(Z[f"c{k}_r{r}_di"]).double()
            gt = torch.from_numpy(Z[f"c{k}_r{r}_dt"]).double()
            assert abs(loss - gl) <= 1e-5 * max(1.0, abs(gl))
            # 3e-5 = the fp32 reference's own distance from the fp64 truth (tests/test_oracle.py); ours is checked at
            # 1e-5 against the fp64 closed form in test_gpu_kernels.py
            assert abs(dls - gd) <= 3e-5 * abs(gd) + 1.2e-7 * c["go"] * max(1.0, c["ls"])
            assert float((torch.from_numpy(di).double() - gi).norm()) <= 1e-5 * float(gi.norm()) + floor
            assert float((torch.from_numpy(dt).double() - gt).norm()) <= 1e-5 * float(gt.norm()) + floor


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("W", [2, 4, 8])
def test_nccl_bf16_tensor_core_sizes_against_oracle(W):
    if _ngpu() < W:
        pytest.skip(f"needs {W} GPUs")
    jobs = []
    for local_loss in (False, True):
        for gwg in (False, True):
            jobs.append(dict(Bl=256, D=512, dtype="bfloat16", seed=77, corr=True, ls=20.0, go=2.0,
                             local_loss=local_loss, gwg=gwg, no_small=True))
            jobs.append(dict(Bl=64, D=512, dtype="bfloat16", seed=177, corr=True, ls=20.0, go=2.0,
                             local_loss=local_loss, gwg=gwg))            # latency path (C2's per-rank shape)
    jobs.append(dict(Bl=256, D=512, dtype="bfloat16", seed=78, corr=True, ls=100.0, go=2.0, local_loss=True, gwg=True,
                     adv=True, no_small=True))
    for local_loss, gwg in ((True, True), (False, False)):
        jobs.append(dict(Bl=256, D=512, dtype="bfloat16", seed=79, corr=True, ls=20.0, go=2.0, local_loss=local_loss, gwg=gwg,
                         graphs=True, no_small=True))
        jobs.append(dict(Bl=64, D=512, dtype="bfloat16", seed=179, corr=True, ls=20.0, go=2.0, local_loss=local_loss, gwg=gwg,
                         graphs=True))                                   # latency path replayed from CUDA graphs
    out = _run(W, jobs)
    for j, job in enumerate(jobs):
        img, txt = _features(W, job)
        ref = O.ref_port_ranks(img.float(), txt.float(), job["ls"], W, job["local_loss"], job["gwg"], grad_output=2.0)
        # saturated softmaxes (the ls = 100 job): true losses / gradients of ~1e-9 sit below the f32 rounding of an
        # LSE of magnitude ls, so those comparisons carry the same absolute floors as the W = 1 tests
        ls_, go_, Bl_ = job["ls"], job["go"], job["Bl"]
        sat = ls_ >= 100.0
        l_floor = 4 * 1.2e-7 * ls_ if sat else 1e-6
        g_floor = 8 * 1.2e-7 * ls_ * go_ * ls_ / (2 * Bl_) * Bl_ ** 0.5 if sat else 0.0
        d_floor = 1.2e-7 * go_ * ls_ * 20 if sat else 1e-7
        for r in range(W):
            loss, di, dt, dls = out[(j, r)]
            assert abs(loss - float(ref[r].loss)) <= 2e-3 * abs(float(ref[r].loss)) + l_floor
            for got, want in ((di, ref[r].d_image), (dt, ref[r].d_text)):
                err = float((torch.from_numpy(got).double() - want.double()).norm())
                assert err <= 2e-3 * float(want.double().norm()) + g_floor
            assert abs(dls - float(ref[r].d_logit_scale)) <= 2e-3 * abs(float(ref[r].d_logit_scale)) + d_floor


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_two_ranks_ragged_shapes_and_forced_fallback():
    """W = 2 on shapes that are NOT tile-aligned (B_l % 128 != 0, D < 512, label offsets that cross tiles): exercises the
    TMA zero-fill tails, row blocks past M in the per-block column partials, merge_col_sums over all columns and the
    [B_g + 2 + B_l] statistics message; plus out-of-window jobs (unnormalised features / ls = 100 with shrunk rows) that
    raise the status word on some ranks, so the predicated one-sided fallback runs at W > 1 on ragged shapes too."""
    W = 2
    jobs = []
    for Bl, D in ((200, 96), (300, 200), (1000, 512), (200, 512), (1000, 96)):
        for local_loss, gwg in ((True, True), (False, False), (True, False)):
            jobs.append(dict(Bl=Bl, D=D, dtype="bfloat16", seed=100 + Bl + D, corr=True, ls=20.0, go=2.0,
                             local_loss=local_loss, gwg=gwg, no_small=True))
    jobs.append(dict(Bl=300, D=200, dtype="bfloat16", seed=5, corr=True, ls=100.0, go=2.0, local_loss=True, gwg=True, adv=True,
                     no_small=True))
    jobs.append(dict(Bl=1000, D=96, dtype="float16", seed=6, corr=True, ls=100.0, go=1.0, local_loss=False, gwg=True, adv=True))
    jobs.append(dict(Bl=200, D=96, dtype="bfloat16", seed=8, corr=True, ls=20.0, go=2.0, local_loss=False, gwg=False))   # latency path, ragged
    for Bl, D, mode in ((300, 200, (True, True)), (1000, 512, (False, False)), (256, 768, (True, True))):
        jobs.append(dict(Bl=Bl, D=D, dtype="bfloat16", seed=300 + Bl, corr=True, ls=20.0, go=2.0, local_loss=mode[0], gwg=mode[1],
                         no_small=True, precopy=True))                   # f16 copies made on the side stream
    out = _run(W, jobs)
    for j, job in enumerate(jobs):
        img, txt = _features(W, job)
        ref = O.ref_port_ranks(img.float(), txt.float(), job["ls"], W, job["local_loss"], job["gwg"], grad_output=job["go"])
        ls_, go_, Bl_ = job["ls"], job["go"], job["Bl"]
        sat = ls_ >= 100.0
        l_floor = 4 * 1.2e-7 * ls_ if sat else 1e-6
        g_floor = 8 * 1.2e-7 * ls_ * go_ * ls_ / (2 * Bl_) * Bl_ ** 0.5 if sat else 0.0
        d_floor = 1.2e-7 * go_ * ls_ * 20 if sat else 1e-7
        for r in range(W):
            loss, di, dt, dls = out[(j, r)]
            assert abs(loss - float(ref[r].loss)) <= 2e-3 * abs(float(ref[r].loss)) + l_floor, (job, r)
            for got, want in ((di, ref[r].d_image), (dt, ref[r].d_text)):
                err = float((torch.from_numpy(got).double() - want.double()).norm())
                assert err <= 2e-3 * float(want.double().norm()) + g_floor, (job, r)
            assert abs(dls - float(ref[r].d_logit_scale)) <= 2e-3 * abs(float(ref[r].d_logit_scale)) + d_floor, (job, r)


def _mismatch_worker(rank, W, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=W, device_id=dev)
    try:
        from mamba_clip_b200 import ClipLoss
        Bl = 64 if rank == 0 else 96          # ranks disagree on B_l
        a = torch.randn(Bl, 64, device=dev, dtype=torch.bfloat16)
        try:
            ClipLoss(True, True, True, rank, W)(a, a, torch.tensor(10.0, device=dev))
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, str(e)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_unequal_shards_raise_on_every_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mismatch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all("same shape" in m for m in res.values()), res


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_two_ranks_loss_from_projections_in_place_gather():
    """Producer epilogue fused into the gather prologue at W = 2: scaled (un-normalised) fp32 projections in, loss and the
    projections' gradients out, against the oracle's rank emulation run through F.normalize in fp64."""
    W, Bl, D, ls, go = 2, 300, 256, 20.0, 2.0
    jobs = [dict(Bl=Bl, D=D, dtype="float32", seed=55, corr=True, ls=ls, go=go, local_loss=ll, gwg=gw, proj=True)
            for ll, gw in ((True, True), (False, False))]
    out = _run(W, jobs)
    import torch.nn.functional as F
    for j, job in enumerate(jobs):
        img, txt = _features(W, job)
        # the worker scales its shard by 3.0 / 0.5 before normalising: normalisation undoes it, the gradient sees 1/scale
        ref = O.ref_port_ranks(F.normalize(img * 3.0, dim=-1).bfloat16().float(), F.normalize(txt * 0.5, dim=-1).bfloat16().float(),
                               ls, W, job["local_loss"], job["gwg"], grad_output=go)
        for r in range(W):
            loss, di, dt, dls = out[(j, r)]
            assert abs(loss - float(ref[r].loss)) <= 2e-3 * abs(float(ref[r].loss)) + 1e-6
            assert abs(dls - float(ref[r].d_logit_scale)) <= 2e-3 * abs(float(ref[r].d_logit_scale)) + 1e-7
            # gradients w.r.t. the raw projections are tangential: check orthogonality to the projection rows and the norm
            # ratio against the feature-gradient of the reference pushed through the fp64 normalisation backward
            for got, want, scale, feats in ((di, ref[r].d_image, 3.0, img), (dt, ref[r].d_text, 0.5, txt)):
                x = (feats[r * Bl:(r + 1) * Bl].double() * scale).requires_grad_(True)
                F.normalize(x, dim=-1).backward(want.double())
                amp = float(want.double().norm()) / max(float((x.grad * float(x.detach().norm(dim=1).mean())).norm()), 1e-30)
                err = float((torch.from_numpy(got).double() - x.grad).norm())
                assert err <= 2e-3 * max(1.0, amp) * float(x.grad.norm()) + 1e-9, (job, r, err / float(x.grad.norm()), amp)
