"""Gradient-accumulation form of the loss (mamba_clip_b200.accum.clip_loss_accum; reference train.py:198-290, intended
semantics train.py:262-270): loss of micro-batch j against the cached features of the other micro-batches == ClipLoss on
the concatenation, with gradients only for micro-batch j.  CPU: host logic on the oracle-emulated primitives; GPU: the
real kernels."""
import pytest
import torch

from oracle import clip_oracle as O


def _chunks(B, D, accum, seed, dtype=torch.float32, correlated=True):
    img, txt = O.make_features(accum * B, D, seed=seed, correlated=correlated, dtype=dtype)
    return img, txt, list(img.split(B)), list(txt.split(B))


@pytest.fixture(params=["one-sided", "two-sided"])
def emulated(request):
    from mamba_clip_b200 import _cabi
    from tests._emul import EmulatedBackend, EmulatedPairBackend
    be = EmulatedPairBackend() if request.param == "two-sided" else EmulatedBackend()
    _cabi.set_backend_override(be)
    yield be
    _cabi.set_backend_override(None)


@pytest.mark.parametrize("accum,B,D,ls", [(2, 16, 32, 10.0), (3, 8, 24, 25.0)])
def test_accum_matches_loss_on_concatenation(emulated, accum, B, D, ls):
    from mamba_clip_b200 import ClipLoss
    from mamba_clip_b200.accum import clip_loss_accum
    img, txt, ci, ct = _chunks(B, D, accum, seed=5, correlated=False)    # unsaturated: gradients well above f32 noise
    ref = O.ref_port_single(img, txt, ls, grad_output=2.0)
    crit = ClipLoss()
    for j in range(accum):
        a = ci[j].clone().requires_grad_(True)
        b = ct[j].clone().requires_grad_(True)
        s = torch.tensor(ls, requires_grad=True)
        out = clip_loss_accum(crit, ci, ct, j, a, b, s)
        assert set(out) == {"contrastive_loss"}
        out["contrastive_loss"].backward(torch.tensor(2.0))
        assert abs(float(out["contrastive_loss"].detach()) - float(ref.loss)) <= 3e-6 * max(1.0, abs(float(ref.loss)))
        lo, hi = j * B, (j + 1) * B
        assert a.grad.shape == (B, D) and O.rel_err(a.grad, ref.d_image[lo:hi]) <= 2e-5
        assert O.rel_err(b.grad, ref.d_text[lo:hi]) <= 2e-5
        assert abs(float(s.grad) - float(ref.d_logit_scale)) <= 3e-5 * abs(float(ref.d_logit_scale)) + 1e-7
    # the backward recompute covered only the live rows: 2 block_grad calls per micro-step
    assert emulated.calls.count("block_grad") == 2 * accum


def test_accum_argument_validation(emulated):
    from mamba_clip_b200 import ClipLoss
    from mamba_clip_b200.accum import clip_loss_accum
    _, _, ci, ct = _chunks(8, 16, 2, seed=1)
    with pytest.raises(ValueError):
        clip_loss_accum(ClipLoss(), ci, ct, 2, ci[0], ct[0], 10.0)
    with pytest.raises(ValueError):
        clip_loss_accum(ClipLoss(), ci, ct[:1], 0, ci[0], ct[0], 10.0)
    with pytest.raises(ValueError):
        clip_loss_accum(ClipLoss(), ci, ct, 0, ci[0][:4], ct[0][:4], 10.0)


@pytest.mark.gpu
@pytest.mark.parametrize("accum,B,D,dtype,tol", [(4, 256, 512, torch.bfloat16, 2e-3), (2, 96, 200, torch.float32, 1e-5)])
def test_accum_on_gpu(accum, B, D, dtype, tol):
    from mamba_clip_b200 import ClipLoss
    from mamba_clip_b200.accum import clip_loss_accum
    ls = 20.0
    # fp32 (1e-5 bar): unsaturated data, so that the gradients sit well above the f32 rounding floor of the LSE
    img, txt, ci, ct = _chunks(B, D, accum, seed=9, dtype=dtype, correlated=dtype != torch.float32)
    ref = O.ref_port_single(img.float(), txt.float(), ls)
    crit = ClipLoss()
    ci_d, ct_d = [c.cuda() for c in ci], [c.cuda() for c in ct]
    for j in (0, accum - 1):
        a = ci_d[j].clone().requires_grad_(True)
        b = ct_d[j].clone().requires_grad_(True)
        s = torch.tensor(ls, device="cuda", requires_grad=True)
        loss = clip_loss_accum(crit, ci_d, ct_d, j, a, b, s, output_dict=False)
        loss.backward()
        lo, hi = j * B, (j + 1) * B
        assert abs(float(loss.detach()) - float(ref.loss)) <= tol * abs(float(ref.loss)) + 3e-6
        assert O.rel_err(a.grad.cpu(), ref.d_image[lo:hi]) <= tol and O.rel_err(b.grad.cpu(), ref.d_text[lo:hi]) <= tol
        assert abs(float(s.grad) - float(ref.d_logit_scale)) <= max(tol, 3e-5) * abs(float(ref.d_logit_scale)) + 1e-7
        # and it equals ClipLoss on the concatenation
        fi = torch.cat(ci_d[:j] + [a.detach()] + ci_d[j + 1:]).requires_grad_(True)
        ft = torch.cat(ct_d[:j] + [b.detach()] + ct_d[j + 1:]).requires_grad_(True)
        s2 = torch.tensor(ls, device="cuda", requires_grad=True)
        l2 = crit(fi, ft, s2, output_dict=False)
        l2.backward()
        # (same forward launches on the general path; sizes that qualify for the latency path sum in another order)
        assert abs(float(l2.detach()) - float(loss.detach())) <= 2e-6 * abs(float(loss.detach()))
        # the recompute launch over a row subset may pick another column split, i.e. another f32 summation order
        assert O.rel_err(a.grad.float().cpu(), fi.grad[lo:hi].float().cpu()) <= 5e-4
        assert O.rel_err(b.grad.float().cpu(), ft.grad[lo:hi].float().cpu()) <= 5e-4
        assert abs(float(s.grad) - float(s2.grad)) <= 1e-4 * abs(float(s2.grad)) + 1e-8
