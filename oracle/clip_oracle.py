"""CPU oracle for the CLIP contrastive-loss hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference algorithm in
``/root/reference/src/mamba_clip/loss.py``.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import it.  The product (``mamba_clip_b200``) never imports anything under ``oracle/``.

Pinning: the reference ships no tests and no golden vectors, so the pin is the
reference itself, imported in the build container: ``tests/golden/make_golden.py`` runs the real
``mamba_clip.loss.ClipLoss`` (single process and 2/4-rank gloo) and commits its outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks every function below against those fixtures.

Three restatements live here:

* :func:`ref_port_single` - line-for-line port of the W=1 path (materialises the logits and
  uses ``F.cross_entropy`` exactly like ``loss.py:109-111,142-145``).  This is also the CPU
  baseline that ``bench.py`` times ("kind": "port").
* :func:`ref_port_ranks` - the W>1 semantics of ``loss.py:16-44,89-113`` emulated in ONE
  process: every rank's loss is built in a single autograd graph so that
  ``gather_with_grad`` (torch's ``_AllGather.backward`` = reduce-scatter SUM) falls out of
  autograd.  Returns the per-rank loss / grads the reference produces on each rank.
* :func:`closed_form` - independent chunked fp64 evaluation of the closed forms in SURVEY.md
  section 3.2; never materialises more than ``chunk x B`` logits, so it reaches the BASELINE
  sizes (B = 32768 / 65536) that the materialising port cannot.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md section 8d / BASELINE.md section 3)
# --------------------------------------------------------------------------------------
def make_features(batch: int, dim: int, seed: int = 1234, correlated: bool = False,
                  dtype: torch.dtype = torch.float32):
    """Unit-normalised image/text features, generated globally on CPU from one seed.

    ``correlated=True`` gives ``T = normalize(I + 0.1*randn)`` (peaky softmax, loss -> 0).
    """
    g = torch.Generator().manual_seed(seed)
    img = F.normalize(torch.randn(batch, dim, generator=g), dim=-1)
    if correlated:
        txt = F.normalize(img + 0.1 * torch.randn(batch, dim, generator=g), dim=-1)
    else:
        txt = F.normalize(torch.randn(batch, dim, generator=g), dim=-1)
    return img.to(dtype), txt.to(dtype)


@dataclass
class RankResult:
    loss: torch.Tensor            # 0-dim
    d_image: torch.Tensor         # [B_l, D]
    d_text: torch.Tensor          # [B_l, D]
    d_logit_scale: torch.Tensor   # 0-dim


# --------------------------------------------------------------------------------------
# 1. faithful W=1 port  (loss.py:109-111 logits, :76-87 labels, :142-145 loss)
# --------------------------------------------------------------------------------------
def ref_port_single(image: torch.Tensor, text: torch.Tensor, logit_scale,
                    grad_output: float = 1.0, need_grad: bool = True) -> RankResult:
    image = image.detach().clone().requires_grad_(need_grad)
    text = text.detach().clone().requires_grad_(need_grad)
    ls = torch.as_tensor(logit_scale, dtype=torch.float32).detach().clone().requires_grad_(need_grad)
    # loss.py:110-111 -- scale is applied to the left operand *before* the GEMM
    logits_per_image = ls * image @ text.T
    logits_per_text = ls * text @ image.T
    # loss.py:79 -- labels = arange(num_logits)
    labels = torch.arange(logits_per_image.shape[0], dtype=torch.long)
    # loss.py:142-145
    loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
    if not need_grad:
        z = torch.zeros(())
        return RankResult(loss.detach(), z, z, z)
    loss.backward(torch.as_tensor(grad_output, dtype=loss.dtype))
    return RankResult(loss.detach(), image.grad, text.grad, ls.grad)


# --------------------------------------------------------------------------------------
# 2. W>1 semantics in one process  (loss.py:16-44 gather, :89-113 logits, :80-81 label offset)
# --------------------------------------------------------------------------------------
def ref_port_ranks(image_all: torch.Tensor, text_all: torch.Tensor, logit_scale, world_size: int,
                   local_loss: bool, gather_with_grad: bool, grad_output: float = 1.0) -> List[RankResult]:
    """Per-rank results of the reference for a global batch split evenly over ``world_size`` ranks.

    Each rank owns a leaf copy of its shard and of ``logit_scale`` (as in DDP).  With
    ``gather_with_grad`` the gathered tensor is the autograd concatenation of all ranks' leaves, so
    d(sum_r loss_r)/d(shard) is what ``_AllGather.backward``'s reduce-scatter(SUM) delivers when
    every rank calls ``backward(grad_output)``.  Without it, the gathered tensor is detached
    except (``not local_loss``) for the rank's own slot (loss.py:37-40).
    """
    W = world_size
    Bg = image_all.shape[0]
    assert Bg % W == 0
    Bl = Bg // W
    img_leaf = [image_all[r * Bl:(r + 1) * Bl].detach().clone().requires_grad_(True) for r in range(W)]
    txt_leaf = [text_all[r * Bl:(r + 1) * Bl].detach().clone().requires_grad_(True) for r in range(W)]
    ls_leaf = [torch.as_tensor(logit_scale, dtype=torch.float32).detach().clone().requires_grad_(True)
               for _ in range(W)]
    losses = []
    for r in range(W):
        if W == 1:
            all_img, all_txt = img_leaf[0], txt_leaf[0]
        elif gather_with_grad:                          # loss.py:25-27
            all_img = torch.cat(img_leaf, dim=0)
            all_txt = torch.cat(txt_leaf, dim=0)
        else:                                           # loss.py:29-42
            gi = [t.detach() for t in img_leaf]
            gt = [t.detach() for t in txt_leaf]
            if not local_loss:
                gi[r] = img_leaf[r]
                gt[r] = txt_leaf[r]
            all_img = torch.cat(gi, dim=0)
            all_txt = torch.cat(gt, dim=0)
        ls = ls_leaf[r]
        if W > 1 and local_loss:                        # loss.py:101-103
            lpi = ls * img_leaf[r] @ all_txt.T
            lpt = ls * txt_leaf[r] @ all_img.T
        elif W > 1:                                     # loss.py:104-108
            lpi = ls * all_img @ all_txt.T
            lpt = lpi.T
        else:                                           # loss.py:110-111
            lpi = ls * all_img @ all_txt.T
            lpt = ls * all_txt @ all_img.T
        n = lpi.shape[0]
        labels = torch.arange(n, dtype=torch.long)
        if W > 1 and local_loss:                        # loss.py:80-81
            labels = labels + n * r
        losses.append((F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2)
    total = sum(losses)
    total.backward(torch.as_tensor(grad_output, dtype=total.dtype))
    out = []
    for r in range(W):
        zi = torch.zeros_like(img_leaf[r])
        out.append(RankResult(losses[r].detach(),
                              img_leaf[r].grad if img_leaf[r].grad is not None else zi,
                              txt_leaf[r].grad if txt_leaf[r].grad is not None else zi.clone(),
                              ls_leaf[r].grad if ls_leaf[r].grad is not None else torch.zeros(())))
    return out


# --------------------------------------------------------------------------------------
# 3. chunked fp64 closed forms - never materialises [B, B]
# --------------------------------------------------------------------------------------
def _lse_rows(a: torch.Tensor, b: torch.Tensor, ls: float, chunk: int):
    """Row log-sum-exp of ls * a @ b.T, chunked over rows of ``a``.  fp64."""
    out = torch.empty(a.shape[0], dtype=torch.float64)
    for s in range(0, a.shape[0], chunk):
        out[s:s + chunk] = torch.logsumexp(ls * (a[s:s + chunk] @ b.T), dim=1)
    return out


def closed_form(image_all: torch.Tensor, text_all: torch.Tensor, logit_scale: float, world_size: int,
                rank: int, local_loss: bool, gather_with_grad: bool, grad_output: float = 1.0,
                chunk: int = 1024, need_grad: bool = True) -> RankResult:
    """Closed forms of SURVEY.md section 3.2 evaluated in fp64 for one rank.

    G = (w_row * P^row + w_col * P^col - w_diag * E) / (2n) on the rank's row block (for dI) and
    column block (for dT); ``n`` and the weights depend on the mode exactly as in the table.
    """
    I = image_all.double()
    T = text_all.double()
    ls = float(logit_scale)
    go = float(grad_output)
    W = world_size
    Bg, D = I.shape
    Bl = Bg // W
    lo, hi = rank * Bl, (rank + 1) * Bl
    row_lse = _lse_rows(I, T, ls, chunk)           # [Bg] LSE_j S_ij
    col_lse = _lse_rows(T, I, ls, chunk)           # [Bg] LSE_i S_ij
    diag = ls * (I * T).sum(dim=1)                 # [Bg] S_ii
    if W == 1 or not local_loss:
        loss = 0.5 * ((row_lse - diag).mean() + (col_lse - diag).mean())
    else:
        loss = 0.5 * ((row_lse[lo:hi] - diag[lo:hi]).mean() + (col_lse[lo:hi] - diag[lo:hi]).mean())
    if not need_grad:
        z = torch.zeros((), dtype=torch.float64)
        return RankResult(loss, z, z, z)

    own_terms_only = W > 1 and local_loss and not gather_with_grad
    n_feat = Bg if (W == 1 or (not local_loss and not gather_with_grad)) else Bl   # 1/(2n) for features
    n_ls = Bl if (W > 1 and local_loss) else Bg                                  # 1/(2n) for d ls
    dI = torch.zeros(Bl, D, dtype=torch.float64)
    dT = torch.zeros(Bl, D, dtype=torch.float64)
    # row block: rows R of S -> dI_r (both softmax terms unless own_terms_only)
    u_sum = 0.0     # sum_{i in R, j} P^row_ij C_ij
    for s in range(lo, hi, chunk):
        e = min(s + chunk, hi)
        C = I[s:e] @ T.T
        S = ls * C
        Prow = torch.exp(S - row_lse[s:e, None])
        u_sum += float((Prow * C).sum())
        G = Prow.clone()
        if not own_terms_only:
            G += torch.exp(S - col_lse[None, :])
        idx = torch.arange(s, e)
        G[idx - s, idx] -= 1.0 if own_terms_only else 2.0
        dI[s - lo:e - lo] = G @ T
    # column block: columns R of S -> dT_r
    v_sum = 0.0     # sum_{i, j in R} P^col_ij C_ij
    for s in range(lo, hi, chunk):
        e = min(s + chunk, hi)
        Ct = T[s:e] @ I.T                          # [chunk, Bg] = C[:, s:e].T
        St = ls * Ct
        Pcol = torch.exp(St - col_lse[s:e, None])
        v_sum += float((Pcol * Ct).sum())
        G = Pcol.clone()
        if not own_terms_only:
            G += torch.exp(St - row_lse[None, :])
        idx = torch.arange(s, e)
        G[idx - s, idx] -= 1.0 if own_terms_only else 2.0
        dT[s - lo:e - lo] = G @ I
    dI *= go * ls / (2.0 * n_feat)
    dT *= go * ls / (2.0 * n_feat)
    c_diag = float((I[lo:hi] * T[lo:hi]).sum())
    t_r = u_sum + v_sum - 2.0 * c_diag
    if W > 1 and local_loss:
        dls = go * t_r / (2.0 * n_ls)
    elif W == 1:
        dls = go * t_r / (2.0 * Bg)
    else:
        # local_loss=False: every rank holds the full sum over all rows/columns
        tot = 0.0
        for r in range(W):
            if r == rank:
                tot += t_r
                continue
            l2, h2 = r * Bl, (r + 1) * Bl
            for s in range(l2, h2, chunk):
                e = min(s + chunk, h2)
                C = I[s:e] @ T.T
                tot += float((torch.exp(ls * C - row_lse[s:e, None]) * C).sum())
                Ct = T[s:e] @ I.T
                tot += float((torch.exp(ls * Ct - col_lse[s:e, None]) * Ct).sum())
            tot -= 2.0 * float((I[l2:h2] * T[l2:h2]).sum())
        dls = go * tot / (2.0 * Bg)
    return RankResult(loss, dI, dT, torch.tensor(dls, dtype=torch.float64))


# --------------------------------------------------------------------------------------
# 3b. the same closed forms split into "global statistics once" + "any subset of rows": what the
#     headline sizes (B = 32768 / 65536) need -- loss and d(logit_scale) are exact over ALL rows and
#     columns, feature gradients are evaluated on a seeded subset of rows against ALL columns.
# --------------------------------------------------------------------------------------
@dataclass
class GlobalStats:
    row_lse: torch.Tensor   # [Bg] fp64  LSE_j S_ij
    col_lse: torch.Tensor   # [Bg] fp64  LSE_i S_ij
    u: torch.Tensor         # [Bg] fp64  sum_j P^row_ij C_ij   (C = raw dots)
    v: torch.Tensor         # [Bg] fp64  sum_i P^col_ij C_ij
    diag: torch.Tensor      # [Bg] fp64  C_ii


def global_stats(image_all: torch.Tensor, text_all: torch.Tensor, logit_scale: float, chunk: int = 1024) -> GlobalStats:
    """Two chunked fp64 passes over S = ls * I T^T (loss.py:102-111 without ever holding [B, B])."""
    I = image_all.double()
    T = text_all.double()
    ls = float(logit_scale)
    Bg = I.shape[0]
    out = [torch.empty(Bg, dtype=torch.float64) for _ in range(4)]
    for k, (A, Bm) in enumerate(((I, T), (T, I))):
        for s in range(0, Bg, chunk):
            C = A[s:s + chunk] @ Bm.T
            S = ls * C
            lse = torch.logsumexp(S, dim=1)
            out[2 * k][s:s + chunk] = lse
            out[2 * k + 1][s:s + chunk] = (torch.exp(S - lse[:, None]) * C).sum(dim=1)
    return GlobalStats(out[0], out[2], out[1], out[3], (I * T).sum(dim=1))


def closed_form_rows(image_all: torch.Tensor, text_all: torch.Tensor, logit_scale: float, world_size: int, rank: int,
                     local_loss: bool, gather_with_grad: bool, stats: GlobalStats, rows: torch.Tensor,
                     grad_output: float = 1.0):
    """Per-rank reference values from precomputed global statistics (same mode table as `closed_form`).

    -> (loss, dI[rows], dT[rows], d_logit_scale); `rows` are LOCAL row indices of the rank's shard.  Loss and
    d(logit_scale) cover every row / column of the rank (or of the global problem for local_loss=False);
    the feature gradients are exact for the requested rows (each needs all B_g columns)."""
    I = image_all.double()
    T = text_all.double()
    ls = float(logit_scale)
    go = float(grad_output)
    W = world_size
    Bg = I.shape[0]
    Bl = Bg // W
    lo, hi = rank * Bl, (rank + 1) * Bl
    st = stats
    per_row = (st.row_lse - ls * st.diag) + (st.col_lse - ls * st.diag)
    if W == 1 or not local_loss:
        loss = 0.5 * per_row.mean()
    else:
        loss = 0.5 * per_row[lo:hi].mean()
    own_terms_only = W > 1 and local_loss and not gather_with_grad
    n_feat = Bg if (W == 1 or (not local_loss and not gather_with_grad)) else Bl
    g = lo + rows.long()
    grads = []
    for A, Bm, lse_x, lse_y in ((I, T, st.row_lse, st.col_lse), (T, I, st.col_lse, st.row_lse)):
        S = ls * (A[g] @ Bm.T)
        G = torch.exp(S - lse_x[g, None])
        if not own_terms_only:
            G += torch.exp(S - lse_y[None, :])
        G[torch.arange(g.numel()), g] -= 1.0 if own_terms_only else 2.0
        grads.append((go * ls / (2.0 * n_feat)) * (G @ Bm))
    t_all = st.u + st.v - 2.0 * st.diag
    if W > 1 and local_loss:
        dls = go * float(t_all[lo:hi].sum()) / (2.0 * Bl)
    else:
        dls = go * float(t_all.sum()) / (2.0 * Bg)
    return loss, grads[0], grads[1], torch.tensor(dls, dtype=torch.float64)


# --------------------------------------------------------------------------------------
# block-level restatements of the two device primitives (used by tests to emulate the C-ABI on
# CPU for the gloo world_size-2 tests of the host logic, and to check the kernels directly)
# --------------------------------------------------------------------------------------
def block_row_lse(x: torch.Tensor, y: torch.Tensor, ls: float, diag_off: Optional[int] = None):
    """lse_i = LSE_j ls*<x_i, y_j>;  diag_i = <x_i, y_{diag_off+i}> (raw dot, 0 if out of range)."""
    C = x.double() @ y.double().T
    lse = torch.logsumexp(ls * C, dim=1)
    diag = torch.zeros(x.shape[0], dtype=torch.float64)
    if diag_off is not None:
        i = torch.arange(x.shape[0])
        j = i + diag_off
        ok = (j >= 0) & (j < y.shape[0])
        diag[ok] = C[i[ok], j[ok]]
    return lse, diag


def block_grad(x: torch.Tensor, y: torch.Tensor, ls: float, lse_x: torch.Tensor,
               lse_y: Optional[torch.Tensor], diag_off: int, w_row: float, w_col: float,
               w_diag: float, alpha: float):
    """dX = alpha * G @ y,  G_ij = w_row*exp(s_ij-lse_x[i]) + w_col*exp(s_ij-lse_y[j]) - w_diag*[j==i+diag_off];
    rowdot_i = sum_j exp(s_ij - lse_x[i]) * <x_i, y_j>."""
    X = x.double()
    Y = y.double()
    C = X @ Y.T
    S = ls * C
    Prow = torch.exp(S - lse_x.double()[:, None])
    G = w_row * Prow
    if w_col != 0.0:
        G = G + w_col * torch.exp(S - lse_y.double()[None, :])
    i = torch.arange(x.shape[0])
    j = i + diag_off
    ok = (j >= 0) & (j < y.shape[0])
    G[i[ok], j[ok]] -= w_diag
    return alpha * (G @ Y), (Prow * C).sum(dim=1)


def rel_err(x: torch.Tensor, ref: torch.Tensor) -> float:
    """||x - ref||_2 / ||ref||_2 (|x-ref|/|ref| for scalars) -- the metric of SURVEY.md section 8c."""
    x = x.detach().double().reshape(-1)
    ref = ref.detach().double().reshape(-1)
    den = float(ref.norm())
    num = float((x - ref).norm())
    if den == 0.0:
        return num
    return num / den
