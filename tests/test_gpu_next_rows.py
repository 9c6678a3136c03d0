"""The "next" rows of the scope table at the same parity bar: the producer epilogue
(F.normalize + cast, reference model.py:1011-1017) and the eval-time contrastive loss (eval.py:107-116)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,D,dtype", [(64, 512, torch.bfloat16), (129, 768, torch.float16), (7, 33, torch.float32),
                                       (4096, 512, torch.bfloat16)])
def test_normalize_features_forward_backward(B, D, dtype):
    from mamba_clip_b200.producer import normalize_features
    g = torch.Generator().manual_seed(B + D)
    x = (torch.randn(B, D, generator=g) * 3.0).cuda().requires_grad_(True)
    y = normalize_features(x, dtype)
    ref = F.normalize(x.detach().double(), dim=-1)
    assert y.dtype == dtype and y.shape == (B, D)
    tol = 1e-6 if dtype == torch.float32 else 2e-3
    assert O.rel_err(y.cpu(), ref.cpu()) <= tol
    go = torch.randn(B, D, generator=g).to(dtype).cuda()
    y.backward(go)
    xr = x.detach().double().requires_grad_(True)
    F.normalize(xr, dim=-1).backward(go.double())
    assert x.grad.dtype == torch.float32
    assert O.rel_err(x.grad.cpu(), xr.grad.cpu()) <= 1e-5


def test_zero_rows_follow_torch_eps_semantics():
    from mamba_clip_b200.producer import normalize_features
    x = torch.zeros(4, 64, device="cuda")
    x[1] = 1.0
    x.requires_grad_(True)
    y = normalize_features(x, torch.float32)
    assert torch.equal(y[0], torch.zeros(64, device="cuda"))           # 0 / max(0, eps) = 0, as F.normalize
    assert abs(float(y[1].norm()) - 1.0) < 1e-6
    y.sum().backward()
    assert torch.isfinite(x.grad).all()


def test_producer_into_loss_chain_matches_fp64_chain():
    """raw fp32 projections -> normalize_features(bf16) -> ClipLoss: gradient w.r.t. the raw projections."""
    from mamba_clip_b200 import ClipLoss
    from mamba_clip_b200.producer import normalize_features
    B, D, ls = 512, 512, 20.0
    g = torch.Generator().manual_seed(3)
    raw_i = torch.randn(B, D, generator=g)
    raw_t = raw_i + 1.0 * torch.randn(B, D, generator=g)      # cos(pair) ~ 0.7: loss ~ 1e-3, well above the f32 floor
    a = raw_i.cuda().requires_grad_(True)
    b = raw_t.cuda().requires_grad_(True)
    loss = ClipLoss()(normalize_features(a), normalize_features(b), torch.tensor(ls, device="cuda"), output_dict=False)
    loss.backward()
    ad = raw_i.double().requires_grad_(True)
    bd = raw_t.double().requires_grad_(True)
    ni, nt = F.normalize(ad, dim=-1), F.normalize(bd, dim=-1)
    # same bf16 rounding of the normalised features as the fused path (straight-through for the gradient)
    ni = ni + (ni.detach().float().bfloat16().double() - ni.detach())
    nt = nt + (nt.detach().float().bfloat16().double() - nt.detach())
    ni.retain_grad()
    nt.retain_grad()
    logits = ls * ni @ nt.T
    lab = torch.arange(B)
    ref = (F.cross_entropy(logits, lab) + F.cross_entropy(logits.T, lab)) / 2
    ref.backward()
    # one f32 ulp of an LSE of magnitude ls is the absolute floor of any loss value
    assert abs(float(loss.detach()) - float(ref)) <= 2e-3 * abs(float(ref)) + 1.2e-7 * ls
    # The loss hands dLoss/dfeature to the producer's backward in bf16 (the input dtype, as the reference's AMP path does):
    # an error of <= 2e-3 of ITS norm.  The normalisation backward then removes the radial component, dx = (g - n <n, g>) / |x|,
    # so the same absolute error is measured against the (smaller) tangential part only.  The amplification
    # A = |g| / |g_tangential| is COMPUTED from the fp64 chain, and the bar is the stated 2e-3 times that factor ...
    # ... plus the usual absolute floor: the LSE travels from forward to backward as one f32 number of magnitude ls, so
    # G = P_row + P_col - 2E carries ~eps*ls absolute noise (tests/test_host_logic.py `grad_floor`), here per unit-norm
    # feature row and divided by the raw projections' norm (~sqrt(D)) by the producer
    floor = 8 * 1.2e-7 * ls * ls / (2 * B) * B ** 0.5 / D ** 0.5
    for got, want, n in ((a.grad, ad.grad, ni), (b.grad, bd.grad, nt)):
        gfull = n.grad
        nd = n.detach()
        gtan = gfull - nd * (nd * gfull).sum(dim=1, keepdim=True)
        amp = float(gfull.norm()) / float(gtan.norm())
        err = float((got.cpu().double() - want).norm())
        print(f"\n[producer chain] amplification |g|/|g_tan| = {amp:.2f}; relative error {err / float(want.norm()):.2e} (bar {2e-3 * amp:.2e})")
        assert amp >= 1.0
        assert err <= 2e-3 * amp * float(want.norm()) + floor


@pytest.mark.parametrize("B,D,dtype,tol", [(64, 512, torch.float32, 1e-5), (256, 512, torch.bfloat16, 2e-3),
                                           (2000, 768, torch.bfloat16, 2e-3)])
def test_eval_contrastive_loss(B, D, dtype, tol):
    from mamba_clip_b200.eval import contrastive_eval_loss
    img, txt = O.make_features(B, D, seed=B, correlated=True, dtype=dtype)
    ls = torch.tensor([14.2857, 14.2857], device="cuda")          # eval.py:106 takes logit_scale.mean()
    out = contrastive_eval_loss(img.cuda(), txt.cuda(), ls)
    ref = O.ref_port_single(img.float(), txt.float(), 14.2857, need_grad=False).loss
    assert out.dim() == 0 and not out.requires_grad
    assert abs(float(out) - float(ref)) <= tol * abs(float(ref)) + 1e-6


@pytest.mark.parametrize("B,D,dtype", [(256, 512, torch.bfloat16), (300, 200, torch.float16), (64, 512, torch.float32)])
def test_clip_loss_from_projections_equals_the_two_step_chain(B, D, dtype):
    """Producer epilogue fused into the gather prologue (`clip_loss_from_projections`) == normalize_features + ClipLoss,
    bit for bit at W = 1 (same kernels, the normalised shard only lands in a different buffer)."""
    from mamba_clip_b200 import ClipLoss
    from mamba_clip_b200.producer import clip_loss_from_projections, normalize_features
    g = torch.Generator().manual_seed(B)
    raw_i = torch.randn(B, D, generator=g) * 2.0
    raw_t = raw_i + 0.7 * torch.randn(B, D, generator=g)
    outs = []
    for fused in (True, False):
        a = raw_i.cuda().requires_grad_(True)
        b = raw_t.cuda().requires_grad_(True)
        s = torch.tensor(20.0, device="cuda", requires_grad=True)
        crit = ClipLoss()
        if fused:
            loss = clip_loss_from_projections(crit, a, b, s, dtype=dtype, output_dict=False)
        else:
            loss = crit(normalize_features(a, dtype), normalize_features(b, dtype), s, output_dict=False)
        loss.backward(torch.tensor(1.5, device="cuda"))
        outs.append((loss.detach(), a.grad, b.grad, s.grad))
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    assert outs[0][1].dtype == torch.float32
