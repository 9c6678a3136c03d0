"""Host-side logic of ClipLoss (mode algebra, gathers, scalar reductions) on CPU: the C-ABI primitives are
replaced by their oracle statements (tests/_emul.py) and the result is compared with the golden vectors
of the real reference, single process and 2/4-rank gloo."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SINGLE = np.load(os.path.join(GOLD, "single.npz"))
RANKS = np.load(os.path.join(GOLD, "ranks.npz"))
SINGLE_CASES = json.loads(str(SINGLE["cases"]))
RANK_CASES = json.loads(str(RANKS["cases"]))


def grad_floor(go, ls, n):
    """Absolute noise floor of a feature gradient: the row/column LSE is carried between forward and
    backward as one fp32 number of magnitude ~ls, so P = exp(s - lse) has ~eps*max(1, ls) relative error;
    times the gradient's natural scale go*ls/(2n)*sqrt(n) (unit-norm rows).  Only matters for saturated
    softmaxes (ls = 100, correlated pairs) where the reference gradient itself is rounding noise."""
    return 4 * 1.2e-7 * max(1.0, ls) * go * ls / (2 * n) * n ** 0.5


@pytest.fixture()
def emulated():
    from mamba_clip_b200 import _cabi
    from tests._emul import EmulatedBackend
    be = EmulatedBackend()
    _cabi.set_backend_override(be)
    yield be
    _cabi.set_backend_override(None)


def test_api_surface_matches_reference():
    import inspect
    import mamba_clip_b200.loss as L
    sig = inspect.signature(L.ClipLoss.__init__)
    assert list(sig.parameters)[1:] == ["local_loss", "gather_with_grad", "cache_labels", "rank", "world_size"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [False, False, False, 0, 1]
    fsig = inspect.signature(L.ClipLoss.forward)
    assert list(fsig.parameters)[1:] == ["image_features", "text_features", "logit_scale", "output_dict", "target"]
    crit = L.ClipLoss()
    assert isinstance(crit, torch.nn.Module) and len(crit.state_dict()) == 0
    for attr in ("local_loss", "gather_with_grad", "cache_labels", "rank", "world_size", "prev_num_logits", "labels"):
        assert hasattr(crit, attr)
    for fn in ("create_loss", "all_gather", "cross_entropy_loss"):
        assert callable(getattr(L, fn))

    class A:
        local_loss, gather_with_grad, rank, world_size = True, True, 3, 8
    c = L.create_loss(A)
    assert (c.local_loss, c.gather_with_grad, c.cache_labels, c.rank, c.world_size) == (True, True, True, 3, 8)
    # labels helper behaves like loss.py:76-87
    lab = c.get_ground_truth(torch.device("cpu"), 4)
    assert lab.tolist() == [12, 13, 14, 15] and c.prev_num_logits == 4


def test_cross_entropy_loss_matches_torch():
    from mamba_clip_b200.loss import cross_entropy_loss
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, 2, generator=g)
    y = torch.randint(0, 2, (16,), generator=g)
    assert torch.allclose(cross_entropy_loss(x, y), torch.nn.functional.cross_entropy(x, y))
    soft = torch.softmax(torch.randn(16, 2, generator=g), -1)
    assert torch.allclose(cross_entropy_loss(x, soft), -(x.log_softmax(-1) * soft).sum(-1).mean())


def test_validation_errors(emulated):
    from mamba_clip_b200 import ClipLoss
    crit = ClipLoss()
    with pytest.raises(ValueError):
        crit(torch.randn(4, 8), torch.randn(5, 8), torch.tensor(1.0))
    with pytest.raises(ValueError):
        crit(torch.randn(4, 8, 2), torch.randn(4, 8, 2), torch.tensor(1.0))
    with pytest.raises(ValueError):
        crit(torch.randn(0, 8), torch.randn(0, 8), torch.tensor(1.0))
    with pytest.raises(RuntimeError):
        ClipLoss(world_size=2)(torch.randn(4, 8), torch.randn(4, 8), torch.tensor(1.0))


def test_cpu_tensors_raise_without_override():
    from mamba_clip_b200 import ClipLoss
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipLoss()(torch.randn(4, 8), torch.randn(4, 8), torch.tensor(1.0))


@pytest.mark.parametrize("k", [k for k, c in enumerate(SINGLE_CASES) if c["B"] <= 129])
def test_single_process_against_golden(emulated, k):
    from mamba_clip_b200 import ClipLoss
    case = SINGLE_CASES[k]
    img, txt = O.make_features(case["B"], case["D"], seed=case["seed"], correlated=case["corr"])
    if case["bf16"]:
        img, txt = img.bfloat16().float(), txt.bfloat16().float()
    img.requires_grad_(True)
    txt.requires_grad_(True)
    ls = torch.tensor(case["ls"], requires_grad=True)
    out = ClipLoss(cache_labels=True)(image_features=img, text_features=txt, logit_scale=ls, target=None)
    assert set(out) == {"contrastive_loss"}
    loss = out["contrastive_loss"]
    loss.backward(torch.tensor(case["go"]))
    gl, gd = float(SINGLE[f"c{k}_loss"]), float(SINGLE[f"c{k}_dls"])
    assert abs(float(loss.detach()) - gl) <= 3e-6 * max(1.0, abs(gl)) + 2e-7
    # LSE is stored in fp32: |lse| ~ ls, so P carries ~eps*ls relative error and d(ls) an absolute floor
    assert abs(float(ls.grad) - gd) <= 3e-5 * abs(gd) + 1.2e-7 * case["go"] * max(1.0, case["ls"])
    floor = grad_floor(case["go"], case["ls"], case["B"])
    for key, g in (("di", img.grad), ("dt", txt.grad)):
        gn = float(SINGLE[f"c{k}_{key}_norm"])
        assert abs(float(g.double().norm()) - gn) <= 2e-5 * gn + floor
        if f"c{k}_{key}_full" in SINGLE:
            ref = torch.from_numpy(SINGLE[f"c{k}_{key}_full"]).double()
            assert float((g.double() - ref).norm()) <= 2e-5 * float(ref.norm()) + floor
    assert emulated.calls.count("row_lse") == 2 and emulated.calls.count("block_grad") == 2


def test_output_dict_false_and_float_scale(emulated):
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(16, 32, seed=3)
    img.requires_grad_(True)
    a = ClipLoss()(img, txt, 10.0, output_dict=False)
    assert a.dim() == 0
    a.backward()
    ref = O.ref_port_single(img.detach(), txt, 10.0)
    assert abs(float(a) - float(ref.loss)) < 1e-5
    assert O.rel_err(img.grad, ref.d_image) < 1e-5


# ---------------------------------------------------------------------------------------------------
# multi-process gloo
# ---------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_case(rank, case):
    from mamba_clip_b200 import ClipLoss, _cabi
    from tests._emul import EmulatedBackend, EmulatedPairBackend, EmulatedSmallBackend
    be = EmulatedSmallBackend() if case.get("small") else (EmulatedPairBackend() if case.get("pair") else EmulatedBackend())
    _cabi.set_backend_override(be)
    W, Bl = case["W"], case["Bl"]
    img, txt = O.make_features(W * Bl, case["D"], seed=case["seed"], correlated=case["corr"])
    if case.get("adv"):
        # identical pairs (cos = 1), then the first half of the image rows shrunk: at ls = 100 the columns of the
        # first half see nothing within ~80 log2 units of the other half's positives -> out of the f32 window
        txt = img.clone()
        img[: W * Bl // 2] *= 0.01
    img = img[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
    txt = txt[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
    ls = torch.tensor(case["ls"], requires_grad=True)
    crit = ClipLoss(case["local_loss"], case["gwg"], True, rank, W)
    loss = crit(img, txt, ls)["contrastive_loss"]
    loss.backward(torch.tensor(case["go"]))
    return (rank, float(loss), img.grad.numpy(), txt.grad.numpy(), float(ls.grad), list(be.calls))


def _worker(rank, world, jobs, port, q):
    """One process per rank runs EVERY job of its world size (a spawn costs ~10 s of interpreter + torch start-up, the jobs
    themselves milliseconds): fresh ClipLoss and fresh emulated backend per job, one process group for all of them."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for key, case in jobs:
            out[key] = _run_case(rank, case)
            dist.barrier()
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _spawn_jobs(world, jobs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, jobs, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    per_rank = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return {key: [per_rank[r][key] for r in range(world)] for key, _ in jobs}


_PAIR_KS = list(range(len(RANK_CASES)))
_SMALL_KS = list(range(len(RANK_CASES)))
_ADV_CASE = dict(W=2, Bl=48, D=512, seed=7, corr=True, ls=100.0, go=2.0, local_loss=True, gwg=True, pair=True, adv=True)
_golden_runs = {}


def _golden_run(variant, k):
    """Results of golden case k through the given host path ("plain" one-sided, "pair" two-sided forward, "small" latency
    path, "adv" the out-of-window case); all jobs of one world size run in a single spawn the first time one is asked for."""
    W = _ADV_CASE["W"] if variant == "adv" else RANK_CASES[k]["W"]
    if W not in _golden_runs:
        jobs = []
        for kk, c in enumerate(RANK_CASES):
            if c["W"] != W:
                continue
            jobs.append((("plain", kk), c))
            if kk in _PAIR_KS:
                jobs.append((("pair", kk), dict(c, pair=True)))
            if kk in _SMALL_KS:
                jobs.append((("small", kk), dict(c, small=True)))
        if W == _ADV_CASE["W"]:
            jobs.append((("adv", 0), _ADV_CASE))
        _golden_runs[W] = _spawn_jobs(W, jobs)
    return _golden_runs[W][(variant, k)]


@pytest.mark.parametrize("k", _PAIR_KS)
def test_gloo_two_sided_forward_against_golden(k):
    """Same golden vectors through the two-sided forward host logic: per-rank column sums gathered and merged, status
    flag clean, predicated fallback calls skipped, text-side `v` taken from the backward launch."""
    case = RANK_CASES[k]
    res = _golden_run("pair", k)
    Bl = case["Bl"]
    floor = grad_floor(case["go"], case["ls"], Bl)
    for rank, loss, di, dt, dls, calls in res:
        gl = float(RANKS[f"c{k}_r{rank}_loss"])
        gd = float(RANKS[f"c{k}_r{rank}_dls"])
        gi = torch.from_numpy(RANKS[f"c{k}_r{rank}_di"]).double()
        gt = torch.from_numpy(RANKS[f"c{k}_r{rank}_dt"]).double()
        assert abs(loss - gl) <= 3e-6 * max(1.0, abs(gl)) + 2e-7
        assert abs(dls - gd) <= 3e-5 * abs(gd) + 1.2e-7 * case["go"] * max(1.0, case["ls"])
        assert float((torch.from_numpy(di).double() - gi).norm()) <= 2e-5 * float(gi.norm()) + floor
        assert float((torch.from_numpy(dt).double() - gt).norm()) <= 2e-5 * float(gt.norm()) + floor
        assert calls.count("pair_lse") == 1 and calls.count("merge_col_sums") == 1
        assert calls.count("row_lse(skipped)") == 2 and calls.count("row_lse") == 0


def test_gloo_two_sided_forward_fallback_on_out_of_window_inputs():
    """ls = 100 and half of the global rows scaled by 0.01: the ranks whose columns fall out of the f32 window must
    take the predicated one-sided path and every rank must still reproduce the reference (oracle rank emulation)."""
    res = _golden_run("adv", 0)
    img, _ = O.make_features(2 * 48, 512, seed=7, correlated=True)
    txt = img.clone()
    img[:48] *= 0.01
    ref = O.ref_port_ranks(img, txt, 100.0, 2, True, True, grad_output=2.0)
    took_fallback = 0
    for rank, loss, di, dt, dls, calls in res:
        took_fallback += calls.count("row_lse") > 0
        assert abs(loss - float(ref[rank].loss)) <= 1e-5 * abs(float(ref[rank].loss)) + 4 * 1.2e-7 * 100.0
        floor = grad_floor(2.0, 100.0, 48)
        assert float((torch.from_numpy(di).double() - ref[rank].d_image.double()).norm()) <= 2e-5 * float(ref[rank].d_image.norm()) + floor
        assert float((torch.from_numpy(dt).double() - ref[rank].d_text.double()).norm()) <= 2e-5 * float(ref[rank].d_text.norm()) + floor
    assert took_fallback >= 1


@pytest.mark.parametrize("k", range(len(RANK_CASES)))
def test_gloo_ranks_against_golden(k):
    case = RANK_CASES[k]
    res = _golden_run("plain", k)
    Bl = case["Bl"]
    floor = grad_floor(case["go"], case["ls"], Bl)
    for rank, loss, di, dt, dls, _calls in res:
        gl = float(RANKS[f"c{k}_r{rank}_loss"])
        gd = float(RANKS[f"c{k}_r{rank}_dls"])
        gi = torch.from_numpy(RANKS[f"c{k}_r{rank}_di"]).double()
        gt = torch.from_numpy(RANKS[f"c{k}_r{rank}_dt"]).double()
        assert abs(loss - gl) <= 3e-6 * max(1.0, abs(gl)) + 2e-7
        assert abs(dls - gd) <= 3e-5 * abs(gd) + 1.2e-7 * case["go"] * max(1.0, case["ls"])
        assert float((torch.from_numpy(di).double() - gi).norm()) <= 2e-5 * float(gi.norm()) + floor
        assert float((torch.from_numpy(dt).double() - gt).norm()) <= 2e-5 * float(gt.norm()) + floor


@pytest.mark.parametrize("k", [0, 3, 7, 12, 20])
def test_single_process_fused_backward_branch_against_reference_golden(k):
    """W = 1 with a backend that offers mclip_fused_grad: both feature gradients from one call, d(logit_scale) through
    Euler's identity (sum of xdot) -- must still reproduce the reference's golden outputs."""
    from mamba_clip_b200 import ClipLoss, _cabi
    from tests._emul import EmulatedFusedBackend
    case = SINGLE_CASES[k % len(SINGLE_CASES)]
    kk = k % len(SINGLE_CASES)
    be = EmulatedFusedBackend()
    _cabi.set_backend_override(be)
    try:
        img, txt = O.make_features(case["B"], case["D"], seed=case["seed"], correlated=case["corr"])
        if case["bf16"]:
            img, txt = img.bfloat16().float(), txt.bfloat16().float()
        a = img.clone().requires_grad_(True)
        b = txt.clone().requires_grad_(True)
        ls = torch.tensor(case["ls"], requires_grad=True)
        loss = ClipLoss()(a, b, ls)["contrastive_loss"]
        loss.backward(torch.tensor(case["go"]))
    finally:
        _cabi.set_backend_override(None)
    assert "fused_grad" in be.calls and "block_grad" not in be.calls
    gl, gd = float(SINGLE[f"c{kk}_loss"]), float(SINGLE[f"c{kk}_dls"])
    assert abs(float(loss) - gl) <= 3e-6 * max(1.0, abs(gl)) + 2e-7
    assert abs(float(ls.grad) - gd) <= 3e-5 * abs(gd) + 4e-7 * case["go"] * max(1.0, case["ls"])
    floor = grad_floor(case["go"], case["ls"], case["B"])
    for key, g in (("di", a.grad), ("dt", b.grad)):
        if f"c{kk}_{key}_full" in SINGLE:
            ref = torch.from_numpy(SINGLE[f"c{kk}_{key}_full"]).double()
            assert float((g.double() - ref).norm()) <= 2e-5 * float(ref.norm()) + floor


def test_double_backward_raises_instead_of_returning_zeros(emulated):
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(8, 16, seed=2)
    a = img.clone().requires_grad_(True)
    loss = ClipLoss()(a, txt, torch.tensor(5.0), output_dict=False)
    (g,) = torch.autograd.grad(loss, a, create_graph=True)
    with pytest.raises(RuntimeError, match="once_differentiable|twice|does not require grad"):
        g.sum().backward()


def test_backend_override_is_refused_without_the_test_opt_in(monkeypatch):
    from mamba_clip_b200 import _cabi
    monkeypatch.delenv("MCLIP_ALLOW_TEST_BACKEND", raising=False)
    with pytest.raises(RuntimeError, match="test hook"):
        _cabi.set_backend_override(object())
    _cabi.set_backend_override(None)


def _mismatch_worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        from mamba_clip_b200 import ClipLoss, _cabi
        from tests._emul import EmulatedBackend
        _cabi.set_backend_override(EmulatedBackend())
        a = torch.randn(8 if rank == 0 else 12, 16)          # the ranks disagree on B_l
        try:
            ClipLoss(True, True, True, rank, 2)(a, a, torch.tensor(10.0))
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, str(e)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_unequal_shards_raise_on_every_rank():
    """SURVEY 8(b) "Errors": equal B_l across ranks is validated before anything is launched (the reference would hang
    or fail inside torch); one tiny all-reduce the first time a (group, shape, dtype) is seen."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_mismatch_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all("same shape" in m for m in res.values()), res


@pytest.mark.parametrize("k", _SMALL_KS)
def test_gloo_latency_path_against_golden(k):
    """The latency path's host logic (B_g <= 1024: ONE packed all-gather, every rank evaluates the whole problem, per-mode
    row ranges / weights / scale factors, no scalar collectives) with the oracle statement of its primitives, against the
    reference's own gloo outputs for all four (local_loss, gather_with_grad) modes at 2 and 4 ranks."""
    case = RANK_CASES[k]
    res = _golden_run("small", k)
    Bl = case["Bl"]
    floor = grad_floor(case["go"], case["ls"], Bl)
    for rank, loss, di, dt, dls, calls in res:
        gl = float(RANKS[f"c{k}_r{rank}_loss"])
        gd = float(RANKS[f"c{k}_r{rank}_dls"])
        gi = torch.from_numpy(RANKS[f"c{k}_r{rank}_di"]).double()
        gt = torch.from_numpy(RANKS[f"c{k}_r{rank}_dt"]).double()
        assert abs(loss - gl) <= 3e-6 * max(1.0, abs(gl)) + 2e-7
        assert abs(dls - gd) <= 3e-5 * abs(gd) + 1.2e-7 * case["go"] * max(1.0, case["ls"])
        assert float((torch.from_numpy(di).double() - gi).norm()) <= 2e-5 * float(gi.norm()) + floor
        assert float((torch.from_numpy(dt).double() - gt).norm()) <= 2e-5 * float(gt.norm()) + floor
        # exactly: pack, forward, backward -- and nothing else from the library
        assert calls == ["small_pack", "small_forward", "small_backward"], calls
