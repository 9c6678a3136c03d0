"""GPU parity tests proper: the CUDA kernels, called through the C ABI, against the CPU oracle on the
same seeded inputs.  fp32 inputs (FFMA path): <= 1e-5 relative.  bf16/f16 inputs (tcgen05 path):
<= 2e-3 relative against the oracle run on the fp32 upcast of the same 16-bit values."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-3, torch.float16: 2e-3}


def backend(path):
    from mamba_clip_b200 import _cabi
    return _cabi.CudaBackend(path=path)


def feats(M, N, D, dtype, seed, correlated=False, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=-1) * scale
    y = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1) * scale
    if correlated:
        k = min(M, N)
        y[:k] = torch.nn.functional.normalize(x[:k] + 0.1 * torch.randn(k, D, generator=g), dim=-1) * scale
    return x.to(dtype), y.to(dtype)


SIMT, TC = 1, 2
ROW_LSE_CASES = [
    # (path, dtype, M, N, D, ls, diag_off)
    (SIMT, torch.float32, 1, 1, 8, 5.0, 0),
    (SIMT, torch.float32, 7, 100, 33, 14.2857, 3),
    (SIMT, torch.float32, 64, 64, 512, 14.2857, 0),
    (SIMT, torch.float32, 129, 257, 512, 100.0, 64),
    (SIMT, torch.float32, 300, 2000, 768, 100.0, 1000),
    (SIMT, torch.bfloat16, 100, 333, 100, 30.0, 7),
    (SIMT, torch.float16, 65, 129, 1032, 30.0, 0),
    (SIMT, torch.float32, 64, 640, 64, -7.5, 0),          # negative scale is legal for a float logit_scale
    (TC, torch.bfloat16, 64, 64, 64, 14.2857, 0),
    (TC, torch.bfloat16, 128, 256, 512, 14.2857, 0),
    (TC, torch.bfloat16, 129, 300, 512, 100.0, 64),
    (TC, torch.bfloat16, 1, 1, 8, 5.0, 0),
    (TC, torch.bfloat16, 200, 1000, 200, 30.0, 400),
    (TC, torch.float16, 300, 1000, 768, 30.0, 17),
    (TC, torch.bfloat16, 512, 4096, 512, 100.0, 1024),
    (TC, torch.bfloat16, 384, 5000, 256, -12.0, 0),
    (TC, torch.bfloat16, 2048, 2048, 512, 14.2857, 0),
]


def _ids(cases):
    return [("simt" if c[0] == SIMT else "tc") + "-" + str(c[1]).split(".")[-1] + "-" + "x".join(str(v) for v in c[2:5])
            for c in cases]


@pytest.mark.parametrize("path,dtype,M,N,D,ls,diag_off", ROW_LSE_CASES, ids=_ids(ROW_LSE_CASES))
def test_row_lse(path, dtype, M, N, D, ls, diag_off):
    be = backend(path)
    x, y = feats(M, N, D, dtype, seed=M * 7 + N, correlated=True)
    xd, yd = x.cuda(), y.cuda()
    lsd = torch.tensor([ls], dtype=torch.float32, device="cuda")
    lse, diag = be.row_lse(xd, yd, lsd, diag_off, True)
    torch.cuda.synchronize()
    ref_lse, ref_diag = O.block_row_lse(x.float(), y.float(), ls, diag_off)
    tol = TOL[dtype]
    # LSE is compared absolutely relative to the logit magnitude it is a statistic of
    assert float((lse.cpu().double() - ref_lse).abs().max()) <= tol * max(1.0, abs(ls)) * 0.5 + 2e-6
    assert float((diag.cpu().double() - ref_diag).abs().max()) <= 2e-6
    lse2, none = be.row_lse(xd, yd, lsd, diag_off, False)
    assert none is None and torch.equal(lse2, lse)       # deterministic, diag optional
    lse3, _, rowdot = be.row_lse(xd, yd, lsd, diag_off, False, True)
    C = x.double() @ y.double().T
    ref_rd = (torch.exp(ls * C - ref_lse[:, None]) * C).sum(dim=1)
    rd_tol = (4 * 1.2e-7 * max(1.0, abs(ls)) + 1e-6) if dtype == torch.float32 else 2e-3
    assert torch.equal(lse3, lse)
    assert float((rowdot.cpu().double() - ref_rd).abs().max()) <= rd_tol * max(1.0, float(ref_rd.abs().max()))


def test_row_lse_strided_rows():
    """leading dimension > D (a column slice of a wider matrix) on both paths."""
    for path, dtype in ((SIMT, torch.float32), (TC, torch.bfloat16)):
        be = backend(path)
        x, y = feats(130, 270, 256, dtype, seed=5)
        xw = torch.zeros(130, 320, dtype=dtype, device="cuda")
        yw = torch.zeros(270, 264, dtype=dtype, device="cuda")
        xw[:, :256] = x.cuda()
        yw[:, :256] = y.cuda()
        lsd = torch.tensor([20.0], device="cuda")
        lse, _ = be.row_lse(xw[:, :256], yw[:, :256], lsd, 0, True)
        ref, _ = O.block_row_lse(x.float(), y.float(), 20.0, 0)
        assert float((lse.cpu().double() - ref).abs().max()) <= TOL[dtype] * 10


GRAD_CASES = [
    # (path, dtype, M, N, D, ls, diag_off, (w_row, w_col, w_diag))
    (SIMT, torch.float32, 1, 1, 8, 5.0, 0, (1, 1, 2)),
    (SIMT, torch.float32, 7, 100, 33, 14.2857, 3, (1, 1, 2)),
    (SIMT, torch.float32, 64, 64, 512, 14.2857, 0, (1, 1, 2)),
    (SIMT, torch.float32, 129, 257, 512, 30.0, 64, (1, 0, 1)),
    (SIMT, torch.float32, 100, 700, 768, 100.0, 300, (1, 1, 2)),
    (SIMT, torch.bfloat16, 100, 333, 100, 30.0, 7, (1, 1, 2)),
    (TC, torch.bfloat16, 64, 64, 64, 14.2857, 0, (1, 1, 2)),
    (TC, torch.bfloat16, 128, 128, 512, 14.2857, 0, (1, 1, 2)),
    (TC, torch.bfloat16, 128, 256, 256, 14.2857, 0, (1, 0, 1)),
    (TC, torch.bfloat16, 129, 300, 512, 30.0, 64, (1, 1, 2)),
    (TC, torch.bfloat16, 1, 1, 8, 5.0, 0, (1, 1, 2)),
    (TC, torch.bfloat16, 200, 1000, 200, 30.0, 400, (1, 1, 2)),
    (TC, torch.float16, 300, 1000, 768, 30.0, 17, (1, 1, 2)),
    (TC, torch.bfloat16, 512, 4096, 512, 100.0, 1024, (1, 1, 2)),
    (TC, torch.bfloat16, 256, 3000, 384, 14.2857, 0, (1, 0, 1)),
    (TC, torch.bfloat16, 2048, 2048, 512, 14.2857, 0, (1, 1, 2)),
    # 512 < D <= 768: CTA-pair kernel with 12 X tiles and ONE S buffer (S one step ahead of dX)
    (TC, torch.bfloat16, 384, 2048, 768, 20.0, 0, (1, 1, 2)),
    (TC, torch.bfloat16, 200, 700, 640, 14.2857, 100, (1, 1, 2)),
    (TC, torch.bfloat16, 130, 300, 520, 30.0, 0, (1, 0, 1)),
    (TC, torch.bfloat16, 1024, 4096, 768, 100.0, 512, (1, 1, 2)),
]


@pytest.mark.parametrize("want_rowdot", [True, False], ids=["rowdot", "norowdot"])
@pytest.mark.parametrize("path,dtype,M,N,D,ls,diag_off,w", GRAD_CASES, ids=_ids(GRAD_CASES))
def test_block_grad(path, dtype, M, N, D, ls, diag_off, w, want_rowdot):
    """want_rowdot=False is what ClipLoss uses (rowdot comes from the forward kernel)."""
    be = backend(path)
    x, y = feats(M, N, D, dtype, seed=M * 3 + N, correlated=True)
    xf, yf = x.float(), y.float()
    lse_x, _ = O.block_row_lse(xf, yf, ls, None)
    lse_y, _ = O.block_row_lse(yf, xf, ls, None)          # column LSE of the same block
    go, inv_2n = 3.0, 1.0 / (2 * M)
    ref_dx, ref_rd = O.block_grad(xf, yf, ls, lse_x, lse_y, diag_off, *w, alpha=go * ls * inv_2n)
    dev = "cuda"
    dx, rd = be.block_grad(x.to(dev), y.to(dev), torch.tensor([ls], device=dev), torch.tensor([go], device=dev),
                           lse_x.float().to(dev), lse_y.float().to(dev) if w[1] else None, diag_off,
                           float(w[0]), float(w[1]), float(w[2]), inv_2n, want_rowdot)
    torch.cuda.synchronize()
    assert dx.dtype == dtype and dx.shape == (M, D)
    tol = TOL[dtype]
    scale = go * abs(ls) * inv_2n * M ** 0.5            # natural size of dX for unit-norm rows
    err = float((dx.cpu().double() - ref_dx).norm())
    assert err <= tol * float(ref_dx.norm()) + 4 * 1.2e-7 * max(1.0, abs(ls)) * scale
    if not want_rowdot:
        assert rd is None
        return
    # rowdot = sum_j P_ij c_ij: P inherits the fp32 rounding of s = ls * c (|s| up to ls), i.e. ~eps * ls relative
    rd_tol = (4 * 1.2e-7 * max(1.0, abs(ls)) + 1e-6) if dtype == torch.float32 else 2e-3
    assert float((rd.cpu().double() - ref_rd).abs().max()) <= rd_tol * max(1.0, float(ref_rd.abs().max()))


PAIR_CASES = [
    # (dtype, M, N, D, ls, diag_off, correlated)
    (torch.bfloat16, 64, 64, 64, 14.2857, 0, True),
    (torch.bfloat16, 1, 1, 8, 5.0, 0, False),
    (torch.bfloat16, 256, 256, 512, 14.2857, 0, True),
    (torch.bfloat16, 129, 300, 512, 30.0, 64, True),
    (torch.bfloat16, 300, 129, 200, 30.0, -20, False),
    (torch.float16, 700, 1000, 256, 25.0, 17, True),
    (torch.bfloat16, 2048, 2048, 512, 14.2857, 0, False),
    (torch.bfloat16, 1000, 5000, 512, 40.0, 1024, True),
    (torch.bfloat16, 384, 640, 128, -12.0, 0, True),     # negative scale
    (torch.bfloat16, 4096, 4096, 512, 100.0, 0, True),   # saturated, all positives near the maximum
    # 512 < D <= 768: k-chunks 8..11 of the X block are streamed through the ring in front of their Y chunk
    (torch.bfloat16, 256, 512, 768, 14.2857, 0, True),
    (torch.float16, 300, 1000, 640, 25.0, 17, True),
    (torch.bfloat16, 2048, 4096, 768, 30.0, 1024, False),
    (torch.bfloat16, 1000, 3000, 520, 20.0, 0, True),
]


@pytest.mark.parametrize("dtype,M,N,D,ls,diag_off,corr", PAIR_CASES,
                         ids=["-".join(str(v).split(".")[-1] for v in c) for c in PAIR_CASES])
def test_pair_lse_two_sided_forward(dtype, M, N, D, ls, diag_off, corr):
    """One pass over S yields row LSE + rowdot AND column LSE (mclip_pair_ref + mclip_pair_lse) == the oracle's
    two one-sided statements; status stays 0 for these well-conditioned inputs; results are deterministic."""
    be = backend(TC)
    x, y = feats(M, N, D, dtype, seed=M * 5 + N, correlated=corr)
    xd, yd = x.cuda(), y.cuda()
    lsd = torch.tensor([ls], dtype=torch.float32, device="cuda")
    assert be.pair_supported(xd, yd)
    diag, ref, status = be.pair_ref(xd, yd, lsd, diag_off)
    row_lse, rowdot, col_lse = be.pair_lse(xd, yd, lsd, ref, status, True, diag=diag, diag_off=diag_off)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    ref_row, ref_diag = O.block_row_lse(x.float(), y.float(), ls, diag_off)
    ref_col, _ = O.block_row_lse(y.float(), x.float(), ls, None)
    tol = TOL[dtype] * max(1.0, abs(ls)) * 0.5 + 2e-6
    assert float((diag.cpu().double() - ref_diag).abs().max()) <= 2e-6
    assert float((row_lse.cpu().double() - ref_row).abs().max()) <= tol
    assert float((col_lse.cpu().double() - ref_col).abs().max()) <= tol
    C = x.double() @ y.double().T
    ref_rd = (torch.exp(ls * C - ref_row[:, None]) * C).sum(dim=1)
    assert float((rowdot.cpu().double() - ref_rd).abs().max()) <= 2e-3 * max(1.0, float(ref_rd.abs().max()))
    # raw column sums (the multi-rank form): ln2 * (ref + log2(sum)) is the same column LSE
    _, _, col_sum = be.pair_lse(xd, yd, lsd, ref, status, False, col_mode=1)
    assert col_sum.numel() == N + 2 and float(col_sum[N]) == float(ref) and col_sum[N + 1:].view(torch.int32).item() == 0
    col2 = be.lse_from_sum(col_sum[:N], ref, status)
    row2, _, col3 = be.pair_lse(xd, yd, lsd, ref, status, False)
    torch.cuda.synchronize()
    assert torch.equal(col2, col_lse) and torch.equal(col3, col_lse) and torch.equal(row2, row_lse)
    assert int(status.item()) == 0


def test_pair_lse_flags_out_of_window_and_predicated_fallback_repairs():
    """Logits spread over more than the f32 window (half of the rows have every logit ~100 nats below the largest
    positive pair): the two-sided pass must raise the status flag, and the predicated one-sided calls chained behind
    it must then produce the right LSEs.  With a clean status the same predicated calls must leave their outputs
    untouched."""
    be = backend(TC)
    B, D, ls = 512, 128, 100.0
    g = torch.Generator().manual_seed(5)
    x = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
    y = x.clone()
    x[: B // 2] *= 0.01                     # these rows: all logits within ~1 nat of 0; the others peak at ls = 100
    x, y = x.bfloat16(), y.bfloat16()
    xd, yd = x.cuda(), y.cuda()
    lsd = torch.tensor([ls], dtype=torch.float32, device="cuda")
    diag, ref, status = be.pair_ref(xd, yd, lsd, 0)
    row_lse, rowdot, col_lse = be.pair_lse(xd, yd, lsd, ref, status, True)
    assert int(status.item()) != 0
    be.row_lse(xd, yd, lsd, 0, False, True, run_if=status, out_lse=row_lse, out_rowdot=rowdot)
    be.row_lse(yd, xd, lsd, 0, False, False, run_if=status, out_lse=col_lse)
    torch.cuda.synchronize()
    ref_row, _ = O.block_row_lse(x.float(), y.float(), ls, 0)
    ref_col, _ = O.block_row_lse(y.float(), x.float(), ls, None)
    assert float((row_lse.cpu().double() - ref_row).abs().max()) <= 2e-3 * ls * 0.5
    assert float((col_lse.cpu().double() - ref_col).abs().max()) <= 2e-3 * ls * 0.5
    # clean status -> predicated calls are no-ops
    status.zero_()
    sentinel = torch.full_like(row_lse, -7.0)
    be.row_lse(xd, yd, lsd, 0, False, False, run_if=status, out_lse=sentinel)
    torch.cuda.synchronize()
    assert bool((sentinel == -7.0).all())
    # and the whole loss is right on such inputs
    from mamba_clip_b200 import ClipLoss
    a = xd.clone().requires_grad_(True)
    b = yd.clone().requires_grad_(True)
    s = torch.tensor(ls, device="cuda", requires_grad=True)
    loss = ClipLoss()(a, b, s, output_dict=False)
    loss.backward()
    want = O.closed_form(x.float(), y.float(), ls, 1, 0, False, False)
    assert abs(float(loss.detach()) - float(want.loss)) <= 2e-3 * abs(float(want.loss)) + 3e-6
    assert abs(float(s.grad) - float(want.d_logit_scale)) <= 2e-3 * abs(float(want.d_logit_scale)) + 1e-4


def test_pair_forward_matches_one_sided_forward_in_the_loss():
    """ClipLoss through the two-sided forward == ClipLoss with it disabled (MCLIP_NO_PAIR_FWD=1), to f32 rounding."""
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(3000, 512, seed=3, correlated=True, dtype=torch.bfloat16)
    outs = []
    for flag in ("0", "1"):
        os.environ["MCLIP_NO_PAIR_FWD"] = flag
        try:
            a = img.cuda().requires_grad_(True)
            b = txt.cuda().requires_grad_(True)
            s = torch.tensor(20.0, device="cuda", requires_grad=True)
            loss = ClipLoss()(a, b, s, output_dict=False)
            loss.backward()
            outs.append((float(loss.detach()), a.grad.float().cpu(), b.grad.float().cpu(), float(s.grad)))
        finally:
            os.environ.pop("MCLIP_NO_PAIR_FWD", None)
    (l0, di0, dt0, ds0), (l1, di1, dt1, ds1) = outs
    assert abs(l0 - l1) <= 1e-5 * abs(l1)      # both are f32 sums of LSEs of magnitude ~ls
    assert abs(ds0 - ds1) <= 1e-4 * abs(ds1) + 1e-8
    assert float((di0 - di1).norm()) <= 1e-3 * float(di1.norm())
    assert float((dt0 - dt1).norm()) <= 1e-3 * float(dt1.norm())


@pytest.mark.parametrize("M,N,D,ls,diag_off,w", [(128, 256, 512, 14.2857, 0, (1, 1, 2)), (300, 1000, 384, 30.0, 17, (1, 1, 2)),
                                                 (640, 768, 512, 20.0, 0, (1, 1, 2)), (512, 4096, 512, 100.0, 1024, (1, 0, 1)),
                                                 (4096, 4096, 512, 14.2857, 0, (1, 1, 2))])
def test_block_grad_persistent_pair_kernel(M, N, D, ls, diag_off, w, persistent_backward):
    """Option bwd_persist = 1: the persistent CTA-pair kernel (pairs walk ranges of (row block, step) units, row blocks cut
    by a range boundary go through f32 partials + fix-up) gives the same dX and rowdot as the oracle.  (640, 768) and
    (4096, 4096) make pairs cross row-block boundaries."""
    be = backend(TC)
    assert be.get_option("bwd_persist") == 1
    x, y = feats(M, N, D, torch.bfloat16, seed=M * 3 + N, correlated=True)
    xf, yf = x.float(), y.float()
    lse_x, _ = O.block_row_lse(xf, yf, ls, None)
    lse_y, _ = O.block_row_lse(yf, xf, ls, None)
    ref_dx, ref_rd = O.block_grad(xf, yf, ls, lse_x, lse_y, diag_off, *w, alpha=3.0 * ls / (2 * M))
    dx, rd = be.block_grad(x.cuda(), y.cuda(), torch.tensor([ls], device="cuda"), torch.tensor([3.0], device="cuda"),
                           lse_x.float().cuda(), lse_y.float().cuda() if w[1] else None, diag_off,
                           float(w[0]), float(w[1]), float(w[2]), 1.0 / (2 * M), True)
    torch.cuda.synchronize()
    scale = 3.0 * ls / (2 * M) * M ** 0.5
    assert float((dx.cpu().double() - ref_dx).norm()) <= 2e-3 * float(ref_dx.norm()) + 8 * 1.2e-7 * max(1.0, ls) * scale
    assert float((rd.cpu().double() - ref_rd).abs().max()) <= 2e-3 * max(1.0, float(ref_rd.abs().max()))


def test_block_grad_shape_selected_persistent_kernel_matches_split_grid():
    """Default option (-1): the launch shape decides.  8192 x 32768 x 512 (a rank of C3 on 4 GPUs: 64 row blocks for 74 pair
    slots) takes the persistent kernel; its dX / rowdot must agree with the split-grid kernel (option 0) on the same inputs
    to summation-order noise, and both with the fp64 oracle on a sample of rows."""
    be = backend(TC)
    old = be.get_option("bwd_persist")
    M, N, D, ls = 8192, 32768, 512, 14.2857
    x, y = feats(M, N, D, torch.bfloat16, seed=77, correlated=False)
    xc, yc = x.cuda(), y.cuda()
    lsd, go = torch.tensor([ls], device="cuda"), torch.tensor([2.0], device="cuda")
    lse_x = be.row_lse(xc, yc, lsd, 0, False)[0]
    lse_y = be.row_lse(yc, xc, lsd, 0, False)[0]
    outs = {}
    try:
        for mode in (-1, 0):
            be.set_option("bwd_persist", mode)
            n0 = be.launch_count()
            dx, rd = be.block_grad(xc, yc, lsd, go, lse_x, lse_y, 0, 1.0, 1.0, 2.0, 1.0 / (2 * N), True)
            torch.cuda.synchronize()
            outs[mode] = (dx.float(), rd.clone(), be.launch_count() - n0)
    finally:
        be.set_option("bwd_persist", old)
    (dx_a, rd_a, _), (dx_s, rd_s, _) = outs[-1], outs[0]
    assert float((dx_a - dx_s).norm()) <= 3e-3 * float(dx_s.norm())          # two bf16 roundings of the same f32 sums
    assert float((rd_a - rd_s).abs().max()) <= 1e-4 * max(1.0, float(rd_s.abs().max()))
    rows = torch.arange(0, M, 61)
    xf, yf = x.float(), y.float()
    alpha = 2.0 * ls / (2 * N)
    ref_dx, _ = O.block_grad(xf[rows], yf, ls, lse_x.cpu()[rows].double(), lse_y.cpu().double(), 0, 1.0, 1.0, 0.0, alpha=alpha)
    ref_dx = ref_dx - alpha * 2.0 * yf[rows].double()          # the positive pair of row r sits in column r
    assert float((dx_a.cpu()[rows].double() - ref_dx).norm()) <= 2e-3 * float(ref_dx.norm())
    assert float((dx_s.cpu()[rows].double() - ref_dx).norm()) <= 2e-3 * float(ref_dx.norm())


@pytest.fixture()
def persistent_backward():
    """Switch the library option for the duration of a test (the environment is only read when the library loads)."""
    be = backend(TC)
    old = be.get_option("bwd_persist")
    be.set_option("bwd_persist", 1)
    yield
    be.set_option("bwd_persist", old)


def load_single():
    z = np.load(os.path.join(GOLD, "single.npz"))
    return z, json.loads(str(z["cases"]))


def projector(dim):
    g = torch.Generator().manual_seed(4242)
    return torch.randn(dim, 8, generator=g, dtype=torch.float64)


Z, CASES = load_single()


@pytest.mark.parametrize("small", [True, False], ids=["latency-path", "general-path"])
@pytest.mark.parametrize("k", range(len(CASES)))
def test_clip_loss_against_reference_golden(k, small, monkeypatch):
    """ClipLoss (W=1) on the GPU vs the outputs of the real reference ClipLoss (tests/golden), once through whatever path
    the size selects (the one-forward-one-backward-kernel latency path for B_g <= 1024) and once with that path switched
    off, so that the general kernels keep their coverage at small shapes."""
    from mamba_clip_b200 import ClipLoss
    if not small:
        monkeypatch.setenv("MCLIP_NO_SMALL_PATH", "1")
    case = CASES[k]
    img, txt = O.make_features(case["B"], case["D"], seed=case["seed"], correlated=case["corr"])
    dtype = torch.bfloat16 if case["bf16"] else torch.float32
    img = img.to(dtype).cuda().requires_grad_(True)
    txt = txt.to(dtype).cuda().requires_grad_(True)
    ls = torch.tensor(case["ls"], device="cuda", requires_grad=True)
    loss = ClipLoss(cache_labels=True)(image_features=img, text_features=txt, logit_scale=ls, target=None)["contrastive_loss"]
    loss.backward(torch.tensor(case["go"], device="cuda"))
    tol = TOL[dtype]
    gl, gd = float(Z[f"c{k}_loss"]), float(Z[f"c{k}_dls"])
    assert loss.dtype == torch.float32 and img.grad.dtype == dtype
    assert abs(float(loss.detach()) - gl) <= tol * max(abs(gl), 1e-3 if case["bf16"] else 1.0) + 3e-6
    d_floor = 1.2e-7 * case["go"] * max(1.0, case["ls"]) * (20 if case["bf16"] else 1)
    # vs the reference's own fp32 output: 3e-5 is the reference's distance from the fp64 truth (tests/test_oracle.py),
    # not ours -- the 1e-5 bar of fp32 inputs is asserted against the fp64 closed form right below
    assert abs(float(ls.grad) - gd) <= max(tol, 3e-5) * abs(gd) + d_floor
    if not case["bf16"]:
        i32, t32 = O.make_features(case["B"], case["D"], seed=case["seed"], correlated=case["corr"])
        cf = O.closed_form(i32, t32, case["ls"], 1, 0, False, False, grad_output=case["go"], need_grad=True)
        assert abs(float(ls.grad) - float(cf.d_logit_scale)) <= 1e-5 * abs(float(cf.d_logit_scale)) + d_floor
    B = case["B"]
    floor = 4 * 1.2e-7 * max(1.0, case["ls"]) * case["go"] * case["ls"] / (2 * B) * B ** 0.5
    if case["bf16"]:
        floor *= 2    # same fp32 exponent arithmetic as the FFMA path, plus ex2.approx (2 ulp)
    for key, g in (("di", img.grad), ("dt", txt.grad)):
        g = g.detach().cpu().double()
        if f"c{k}_{key}_full" in Z:
            ref = torch.from_numpy(Z[f"c{k}_{key}_full"]).double()
            assert float((g - ref).norm()) <= tol * float(ref.norm()) + floor
        else:
            ref = torch.from_numpy(Z[f"c{k}_{key}_proj"])
            assert float((g @ projector(g.shape[1]) - ref).norm()) <= tol * float(ref.norm()) + 8 * floor


@pytest.mark.parametrize("B,D,dtype,ls", [(4096, 512, torch.bfloat16, 14.2857), (4096, 512, torch.bfloat16, 100.0),
                                          (3000, 768, torch.float16, 30.0), (1024, 512, torch.float32, 100.0)])
def test_clip_loss_medium_against_closed_form(B, D, dtype, ls):
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(B, D, seed=99, correlated=True, dtype=dtype)
    ref = O.closed_form(img.float(), txt.float(), ls, 1, 0, False, False, grad_output=2.0)
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(ls, device="cuda", requires_grad=True)
    loss = ClipLoss()(a, b, s, output_dict=False)
    loss.backward(torch.tensor(2.0, device="cuda"))
    tol = TOL[dtype]
    assert abs(float(loss.detach()) - float(ref.loss)) <= tol * abs(float(ref.loss)) + 3e-6
    # absolute floor for saturated softmaxes (ls = 100, correlated): s = ls*c is rounded in fp32 (|s| ~ ls), so
    # G = P_row + P_col - 2E carries ~eps*ls absolute noise; times the gradient's natural scale go*ls/(2B)*sqrt(B)
    floor = 8 * 1.2e-7 * max(1.0, ls) * 2.0 * ls / (2 * B) * B ** 0.5
    for got, want in ((a.grad, ref.d_image), (b.grad, ref.d_text)):
        assert float((got.cpu().double() - want).norm()) <= tol * float(want.norm()) + floor
    assert abs(float(s.grad) - float(ref.d_logit_scale)) <= tol * abs(float(ref.d_logit_scale)) + \
        1.2e-7 * 2.0 * max(1.0, ls) * 4
    print(f"\n[parity B={B} D={D} {dtype} ls={ls}] unfloored relative errors: loss "
          f"{abs(float(loss.detach()) - float(ref.loss)) / abs(float(ref.loss)):.2e}  dI {O.rel_err(a.grad.cpu(), ref.d_image):.2e}  "
          f"dT {O.rel_err(b.grad.cpu(), ref.d_text):.2e}  dls {abs(float(s.grad) - float(ref.d_logit_scale)) / abs(float(ref.d_logit_scale)):.2e}")


def test_symmetry_and_homogeneity_properties():
    """Size-independent properties: swapping image/text swaps the gradients; Euler homogeneity
    sum_i <I_i, dI_i> = sum_j <T_j, dT_j> = ls * d(ls)."""
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(1536, 512, seed=11, correlated=True, dtype=torch.bfloat16)
    outs = []
    for a0, b0 in ((img, txt), (txt, img)):
        a = a0.cuda().requires_grad_(True)
        b = b0.cuda().requires_grad_(True)
        s = torch.tensor(25.0, device="cuda", requires_grad=True)
        loss = ClipLoss()(a, b, s, output_dict=False)
        loss.backward()
        outs.append((float(loss.detach()), a.grad.float(), b.grad.float(), float(s.grad), a.detach().float(), b.detach().float()))
    (l0, di0, dt0, ds0, a0, b0), (l1, di1, dt1, ds1, _, _) = outs
    assert abs(l0 - l1) <= 1e-6 * abs(l0) and abs(ds0 - ds1) <= 1e-5 * abs(ds0) + 1e-9
    # the two-sided forward sums rows and columns in different orders, so the swapped run sees LSEs that differ in
    # the last f32 bit: the bf16 gradients agree up to rare one-ulp flips rather than bit for bit
    assert float((di0 - dt1).norm()) <= 5e-4 * float(dt1.norm()) and float((dt0 - di1).norm()) <= 5e-4 * float(di1.norm())
    e_i = float((a0 * di0).sum())
    e_t = float((b0 * dt0).sum())
    assert abs(e_i - 25.0 * ds0) <= 4e-3 * abs(25.0 * ds0) + 1e-6
    assert abs(e_t - 25.0 * ds0) <= 4e-3 * abs(25.0 * ds0) + 1e-6


def test_autocast_rounds_fp32_features_like_the_reference_matmul():
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(256, 512, seed=21, correlated=True)
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(14.2857, device="cuda", requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = ClipLoss()(a, b, s)["contrastive_loss"]
    loss.backward()
    assert loss.dtype == torch.float32 and a.grad.dtype == torch.float32
    ref = O.closed_form(img.bfloat16().float(), txt.bfloat16().float(), 14.2857, 1, 0, False, False)
    assert abs(float(loss.detach()) - float(ref.loss)) <= 2e-3 * abs(float(ref.loss))
    assert O.rel_err(a.grad.cpu(), ref.d_image) <= 2e-3


def test_grad_only_where_required():
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(64, 64, seed=1)
    a = img.cuda().requires_grad_(True)
    loss = ClipLoss()(a, txt.cuda(), 10.0, output_dict=False)
    loss.backward()
    assert a.grad is not None
    with torch.no_grad():
        l2 = ClipLoss()(img.cuda(), txt.cuda(), torch.tensor(10.0, device="cuda"), output_dict=False)
    assert abs(float(l2) - float(loss.detach())) < 1e-6


def test_error_paths_on_gpu():
    from mamba_clip_b200 import ClipLoss, _cabi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipLoss()(torch.randn(4, 8), torch.randn(4, 8), torch.tensor(1.0))
    be = _cabi.CudaBackend(path=2)
    x = torch.randn(8, 12, device="cuda")
    with pytest.raises(RuntimeError, match="tcgen05 path cannot run"):
        be.row_lse(x, x, torch.ones(1, device="cuda"), 0, False)   # fp32 cannot be forced onto tensor cores


def test_launch_counter_moves():
    from mamba_clip_b200 import ClipLoss, _cabi
    be = _cabi.get_backend()
    n0 = be.launch_count()
    img, txt = O.make_features(64, 64, seed=1, dtype=torch.bfloat16)
    a = img.cuda().requires_grad_(True)
    ClipLoss()(a, txt.cuda(), torch.tensor(10.0, device="cuda"), output_dict=False).backward()
    assert be.launch_count() - n0 == 2          # latency path: one forward kernel, one backward kernel


def test_general_path_launch_counter(monkeypatch):
    from mamba_clip_b200 import ClipLoss, _cabi
    monkeypatch.setenv("MCLIP_NO_SMALL_PATH", "1")
    be = _cabi.get_backend()
    n0 = be.launch_count()
    img, txt = O.make_features(64, 64, seed=1, dtype=torch.bfloat16)
    a = img.cuda().requires_grad_(True)
    ClipLoss()(a, txt.cuda(), torch.tensor(10.0, device="cuda"), output_dict=False).backward()
    assert be.launch_count() - n0 >= 6


# ---------------------------------------------------------------------------------------------------------------------
# shared-recompute backward (mclip_fused_grad): dX, dY and xdot from ONE recompute of S
# ---------------------------------------------------------------------------------------------------------------------
FUSED_CASES = [
    # (dtype, M, N, D, ls, diag_off)
    (torch.bfloat16, 1, 1, 8, 5.0, 0),
    (torch.bfloat16, 64, 64, 64, 14.2857, 0),
    (torch.bfloat16, 129, 300, 512, 30.0, 64),
    (torch.bfloat16, 200, 1000, 200, 30.0, 400),
    (torch.float16, 300, 1000, 384, 30.0, 17),
    (torch.bfloat16, 512, 4096, 512, 100.0, 1024),
    (torch.bfloat16, 2048, 2048, 512, 14.2857, 0),
    (torch.bfloat16, 2100, 8200, 256, 20.0, 0),          # several row panels, ragged everything
    (torch.bfloat16, 5000, 8192, 512, 14.2857, 100),     # panel boundary inside the row range
]


@pytest.mark.parametrize("dtype,M,N,D,ls,diag_off", FUSED_CASES,
                         ids=[str(c[0]).split(".")[-1] + "-" + "x".join(str(v) for v in c[1:4]) for c in FUSED_CASES])
def test_fused_grad(dtype, M, N, D, ls, diag_off):
    be = backend(TC)
    x, y = feats(M, N, D, dtype, seed=M * 5 + N, correlated=True)
    xf, yf = x.float(), y.float()
    lse_x, _ = O.block_row_lse(xf, yf, ls, None)
    lse_y, _ = O.block_row_lse(yf, xf, ls, None)
    go, inv_2n = 3.0, 1.0 / (2 * M)
    alpha = go * ls * inv_2n
    # oracle: G once, both products (fp64)
    C = xf.double() @ yf.double().T
    S = ls * C
    G = torch.exp(S - lse_x[:, None]) + torch.exp(S - lse_y[None, :])
    i = torch.arange(M)
    j = i + diag_off
    ok = (j >= 0) & (j < N)
    G[i[ok], j[ok]] -= 2.0
    ref_dx, ref_dy, ref_xdot = alpha * (G @ yf.double()), alpha * (G.T @ xf.double()), (G * C).sum(dim=1)
    dx, dy, xdot = be.fused_grad(x.cuda(), y.cuda(), torch.tensor([ls], device="cuda"), torch.tensor([go], device="cuda"),
                                 lse_x.float().cuda(), lse_y.float().cuda(), diag_off, inv_2n)
    torch.cuda.synchronize()
    tol = TOL[dtype]
    scale = alpha * M ** 0.5
    floor = 8 * 1.2e-7 * max(1.0, ls) * scale
    assert float((dx.cpu().double() - ref_dx).norm()) <= tol * float(ref_dx.norm()) + floor
    assert float((dy.cpu().double() - ref_dy).norm()) <= tol * float(ref_dy.norm()) + floor * (N / M) ** 0.5
    # t = sum_i xdot_i is what d(logit_scale) is made of; per-row values within the bf16 bar of the row's scale
    assert abs(float(xdot.double().sum().cpu()) - float(ref_xdot.sum())) <= tol * max(float(ref_xdot.abs().sum()), 1e-6) + 1e-5
    # determinism: the dY accumulation order is fixed (single owner per tile and launch, stream-ordered panels)
    dx2, dy2, xdot2 = be.fused_grad(x.cuda(), y.cuda(), torch.tensor([ls], device="cuda"), torch.tensor([go], device="cuda"),
                                    lse_x.float().cuda(), lse_y.float().cuda(), diag_off, inv_2n)
    assert torch.equal(dx, dx2) and torch.equal(dy, dy2) and torch.equal(xdot, xdot2)


def test_fused_backward_matches_two_launch_backward():
    """ClipLoss at a size where the shared-recompute backward is chosen (W = 1, B >= 16384) against the same loss with the
    option switched off (two recompute launches): same loss, gradients within bf16 rounding of each other."""
    from mamba_clip_b200 import ClipLoss, _cabi
    be = _cabi.get_backend()
    img, txt = O.make_features(16384, 512, seed=17, dtype=torch.bfloat16)
    assert be.fused_supported(img.cuda(), txt.cuda())
    outs = []
    for fused in (1, 0):
        be.set_option("fused_bwd", fused)
        try:
            a = img.cuda().requires_grad_(True)
            b = txt.cuda().requires_grad_(True)
            s = torch.tensor(14.2857, device="cuda", requires_grad=True)
            n0 = be.launch_count()
            loss = ClipLoss()(a, b, s, output_dict=False)
            loss.backward()
            torch.cuda.synchronize()
            outs.append((float(loss.detach()), a.grad.float(), b.grad.float(), float(s.grad), be.launch_count() - n0))
        finally:
            be.set_option("fused_bwd", 1)
    (l1, di1, dt1, ds1, _), (l0, di0, dt0, ds0, _) = outs
    assert l1 == l0
    assert float((di1 - di0).norm()) <= 2e-3 * float(di0.norm()) and float((dt1 - dt0).norm()) <= 2e-3 * float(dt0.norm())
    assert abs(ds1 - ds0) <= 2e-3 * abs(ds0)


# ---------------------------------------------------------------------------------------------------------------------
# parity AT THE BENCHMARKED SIZES (BASELINE.json configs C3 and the D = 768 shape of C4), against the chunked fp64
# closed form: loss and d(logit_scale) over all rows / columns, dI and dT on 512 seeded rows against all columns
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,D,ls,corr", [(32768, 512, 14.2857, False), (32768, 512, 100.0, True), (16384, 768, 14.2857, False)],
                         ids=["C3-ls14", "C3-ls100-correlated", "D768-16k"])
def test_clip_loss_at_headline_size_against_closed_form(B, D, ls, corr):
    from mamba_clip_b200 import ClipLoss
    img, txt = O.make_features(B, D, seed=1234, correlated=corr, dtype=torch.bfloat16)
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.global_stats(img.float(), txt.float(), ls)
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(99))[:512].sort().values
    go = 2.0
    r_loss, r_di, r_dt, r_dls = O.closed_form_rows(img.float(), txt.float(), ls, 1, 0, True, True, st, rows, grad_output=go)
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(ls, device="cuda", requires_grad=True)
    loss = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)(a, b, s)["contrastive_loss"]
    loss.backward(torch.tensor(go, device="cuda"))
    torch.cuda.synchronize()
    tol = 2e-3
    sat = ls >= 100.0 and corr          # saturated softmax: the true loss / gradients are ~eps-sized, floors as at W = 1
    l_floor = 4 * 1.2e-7 * ls if sat else 0.0
    g_floor = 8 * 1.2e-7 * ls * go * ls / (2 * B) * 512 ** 0.5 if sat else 0.0
    d_floor = 1.2e-7 * go * ls * 20 if sat else 0.0
    e_loss = abs(float(loss.detach()) - float(r_loss))
    e_di = float((a.grad[rows.cuda()].double().cpu() - r_di).norm())
    e_dt = float((b.grad[rows.cuda()].double().cpu() - r_dt).norm())
    e_dls = abs(float(s.grad) - float(r_dls))
    print(f"\n[parity B={B} D={D} ls={ls}] unfloored relative errors: loss {e_loss / abs(float(r_loss)):.2e}  "
          f"dI {e_di / float(r_di.norm()):.2e}  dT {e_dt / float(r_dt.norm()):.2e}  dls {e_dls / abs(float(r_dls)):.2e}")
    assert e_loss <= tol * abs(float(r_loss)) + l_floor
    assert e_di <= tol * float(r_di.norm()) + g_floor
    assert e_dt <= tol * float(r_dt.norm()) + g_floor
    assert e_dls <= tol * abs(float(r_dls)) + d_floor


# ---------------------------------------------------------------------------------------------------------------------
# latency path (mclip_small_forward / mclip_small_backward): the primitives against the fp64 oracle
# ---------------------------------------------------------------------------------------------------------------------
SMALL_CASES = [
    # (dtype, W, Bl, D, ls, (local_loss, gwg))
    (torch.float32, 1, 64, 512, 14.2857, (False, False)),
    (torch.float32, 1, 7, 8, 5.0, (False, False)),
    (torch.bfloat16, 1, 256, 512, 100.0, (False, False)),
    (torch.bfloat16, 8, 64, 512, 14.2857, (False, False)),       # C2
    (torch.bfloat16, 8, 64, 512, 20.0, (True, True)),
    (torch.float16, 4, 100, 96, 30.0, (True, False)),
    (torch.float32, 2, 130, 200, 20.0, (False, True)),
    (torch.bfloat16, 16, 64, 256, 14.2857, (False, False)),
]


@pytest.mark.parametrize("dtype,W,Bl,D,ls,mode", SMALL_CASES,
                         ids=[f"{str(c[0]).split('.')[-1]}-W{c[1]}-{c[2]}x{c[3]}-{int(c[5][0])}{int(c[5][1])}" for c in SMALL_CASES])
def test_small_path_primitives_every_rank(dtype, W, Bl, D, ls, mode):
    """One process plays every rank in turn: the blocked layout [W][2][Bl][D] is what the all-gather of the packed shards
    produces; loss, dA, dB, d(logit_scale) of each rank against the oracle's rank emulation (the reference's semantics)."""
    be = backend(0)
    local_loss, gwg = mode
    Bg = W * Bl
    assert be.small_supported(Bl, Bg, D, dtype)
    img, txt = O.make_features(Bg, D, seed=7 * W + Bl, correlated=True, dtype=dtype)
    ref = O.ref_port_ranks(img.float(), txt.float(), ls, W, local_loss, gwg, grad_output=2.0)
    lsd = torch.tensor([ls], device="cuda")
    go = torch.tensor([2.0], device="cuda")
    if W > 1:
        recv = torch.stack([torch.stack((img[r * Bl:(r + 1) * Bl], txt[r * Bl:(r + 1) * Bl])) for r in range(W)]).cuda()
        A, Bm, stride = recv, recv[0, 1], 2 * Bl * D
    else:
        A, Bm, stride = img.cuda(), txt.cuda(), Bl * D
    tol = TOL[dtype]
    own_only = W > 1 and local_loss and not gwg
    w = (1.0, 0.0, 1.0) if own_only else (1.0, 1.0, 2.0)
    n_feat = Bg if (W == 1 or (not local_loss and not gwg)) else Bl
    n_ls = Bl if (W > 1 and local_loss) else Bg
    for r in range(W):
        off = r * Bl
        lo, hi = (off, off + Bl) if (W > 1 and local_loss) else (0, Bg)
        stats = be.small_forward(A, Bm, Bl, Bg, D, stride, lsd, lo, hi)
        dA, dB, dls = be.small_backward(A, Bm, Bl, Bg, D, stride, lsd, go, stats, off, *w, 1.0 / (2 * n_feat), 1.0 / (2 * n_ls))
        torch.cuda.synchronize()
        loss = float(stats[5 * Bg])
        sat = ls >= 100.0
        assert abs(loss - float(ref[r].loss)) <= tol * abs(float(ref[r].loss)) + (4 * 1.2e-7 * ls if sat else 2e-6)
        floor = 8 * 1.2e-7 * max(1.0, ls) * 2.0 * ls / (2 * n_feat) * Bl ** 0.5
        for got, want in ((dA, ref[r].d_image), (dB, ref[r].d_text)):
            assert float((got.cpu().double() - want.double()).norm()) <= tol * float(want.double().norm()) + floor
        d_floor = 1.2e-7 * 2.0 * max(1.0, ls) * (20 if dtype != torch.float32 else 4)
        assert abs(float(dls) - float(ref[r].d_logit_scale)) <= tol * abs(float(ref[r].d_logit_scale)) + d_floor
    # the launches leave the counter words zeroed: a second identical call reproduces the result bit for bit
    stats2 = be.small_forward(A, Bm, Bl, Bg, D, stride, lsd, lo, hi)
    assert torch.equal(stats, stats2)


def test_small_pack_concatenates_and_casts():
    be = backend(0)
    a = torch.randn(64, 512, device="cuda")
    b = torch.randn(64, 512, device="cuda")
    out = be.small_pack(a, b, torch.bfloat16)
    assert out.shape == (2, 64, 512) and torch.equal(out[0], a.bfloat16()) and torch.equal(out[1], b.bfloat16())


def test_unnormalised_bf16_features_inside_the_f16_range_keep_exact_gradients():
    """the bf16 backward multiplies G with an f16 copy of the features.  Inside the f16 range (here rows of norm
    ~40, elements up to ~6) nothing is lost; the documented limit is |v| <= 65504 (loss.py / INTEGRATION.md)."""
    from mamba_clip_b200 import ClipLoss
    B, D, ls = 2048, 256, 0.01            # un-normalised rows: keep the logits moderate through a small scale
    g = torch.Generator().manual_seed(3)
    img = (torch.randn(B, D, generator=g) * 2.5).bfloat16()
    txt = (img.float() + 0.5 * torch.randn(B, D, generator=g)).bfloat16()
    ref = O.closed_form(img.float(), txt.float(), ls, 1, 0, False, False, grad_output=1.0)
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(ls, device="cuda", requires_grad=True)
    loss = ClipLoss()(a, b, s, output_dict=False)
    loss.backward()
    assert abs(float(loss.detach()) - float(ref.loss)) <= 2e-3 * abs(float(ref.loss)) + 1e-6
    assert O.rel_err(a.grad.cpu(), ref.d_image) <= 2e-3 and O.rel_err(b.grad.cpu(), ref.d_text) <= 2e-3
    assert abs(float(s.grad) - float(ref.d_logit_scale)) <= 2e-3 * abs(float(ref.d_logit_scale))


@pytest.mark.parametrize("M,N,D", [(300, 1000, 512), (256, 2048, 768)])
def test_block_grad_with_a_preconverted_f16_copy_is_bit_identical(M, N, D):
    """mclip_block_grad(Y16 = mclip_convert_f16(Y)) == mclip_block_grad with its own conversion pass (bf16 inputs)."""
    be = backend(TC)
    x, y = feats(M, N, D, torch.bfloat16, seed=M + N, correlated=True)
    lse_x, _ = O.block_row_lse(x.float(), y.float(), 20.0, None)
    lse_y, _ = O.block_row_lse(y.float(), x.float(), 20.0, None)
    xd, yd = x.cuda(), y.cuda()
    args = (xd, yd, torch.tensor([20.0], device="cuda"), torch.tensor([1.0], device="cuda"), lse_x.float().cuda(),
            lse_y.float().cuda(), 0, 1.0, 1.0, 2.0, 0.5 / M, True)
    y16 = be.to_f16(yd)
    # exact for normal f16 values; elements below 6.1e-5 land on the f16 subnormal grid (spacing 2^-24)
    assert y16.dtype == torch.float16 and float((y16.float() - yd.float()).abs().max()) <= 2.0 ** -25
    dx0, rd0 = be.block_grad(*args)
    dx1, rd1 = be.block_grad(*args, y16=y16)
    assert torch.equal(dx0, dx1) and torch.equal(rd0, rd1)
