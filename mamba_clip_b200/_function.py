"""Autograd Function + collective plumbing behind `ClipLoss.forward`.

The reference builds this with implicit autograd over `logit_scale * a @ b.T` and
`F.cross_entropy` (reference src/mamba_clip/loss.py:89-113,142-145) and torch's `_AllGather`.
Here the forward is two calls of the fused row-LSE kernel and the backward two calls of the fused
recompute kernel; see DESIGN.md for the decomposition:

  rank r owns rows R of S = ls * I T^T (block "I_r x all T") and columns R of S (block "T_r x all I").
  forward : row_lse[R] from the first block, col_lse[R] + diag from the second -> loss_r
  exchange: features are all-gathered once (NCCL, contiguous buffers); for the modes whose gradient has
            cross terms the two O(B) LSE vectors are all-gathered as well (8*B_l bytes per rank)
  backward: dI_r and dT_r are each complete on the owning rank -> no gradient collective at all; only
            `local_loss=False` needs one scalar all-reduce for d(logit_scale).
Per-rank values reproduce the reference exactly, including the W x factors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi

_SUPPORTED = (torch.float32, torch.bfloat16, torch.float16)


def _gather_rows_async(x: torch.Tensor, world_size: int, group):
    """Start an all-gather of `[B_l, ...]` shards into one contiguous `[W*B_l, ...]` buffer (no list + cat copy).
    Returns (buffer, work); `work.wait()` makes the current stream wait for the collective."""
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    try:
        work = dist.all_gather_into_tensor(out, x.contiguous(), group=group, async_op=True)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        work = dist.all_gather(list(out.chunk(world_size, dim=0)), x.contiguous(), group=group, async_op=True)
    return out, work


def _gather_rows(x: torch.Tensor, world_size: int, group) -> torch.Tensor:
    out, work = _gather_rows_async(x, world_size, group)
    work.wait()
    return out


def _compute_dtype(t: torch.Tensor) -> torch.dtype:
    """dtype the kernels run in.  Inside an autocast region fp32 features are rounded to the autocast
    dtype, which is what the reference's `@` does there; otherwise the input dtype."""
    if t.dtype == torch.float32 and t.is_cuda and torch.is_autocast_enabled():
        ac = torch.get_autocast_dtype("cuda")
        if ac in (torch.bfloat16, torch.float16):
            return ac
    if t.dtype in _SUPPORTED:
        return t.dtype
    return torch.float32


class ClipLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, local_loss, gather_with_grad, rank, world_size,
                group):
        be = _cabi.get_backend()
        dev = image_features.device
        cdt = _compute_dtype(image_features)
        xi = image_features.detach().to(cdt).contiguous()
        xt = text_features.detach().to(cdt).contiguous()
        if torch.is_tensor(logit_scale):
            ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        else:
            ls = torch.full((1,), float(logit_scale), dtype=torch.float32, device=dev)
        W = int(world_size)
        Bl = xi.shape[0]
        if W > 1:
            # both gathers are in flight at once; the image gather overlaps the first kernel
            all_t, work_t = _gather_rows_async(xt, W, group)
            all_i, work_i = _gather_rows_async(xi, W, group)
            off = int(rank) * Bl
            work_t.wait()
        else:
            all_i, all_t, off = xi, xt, 0
            work_i = None

        # u, v: softmax-weighted raw dots (sum_j P_ij <x_i, y_j>) of the two blocks -- all d(logit_scale) needs
        need_ls = torch.is_tensor(logit_scale) and logit_scale.requires_grad
        if getattr(be, "pair_supported", lambda *_: False)(xi, all_t):
            # two-sided forward: one pass over the rank's row block S_r = ls * I_r T_all^T gives its row LSEs and the
            # sums of every column over its rows (1 GEMM unit instead of 2).  At W > 1 the per-rank column sums are
            # all-gathered (4 * (B_g + 2) bytes per rank) and each rank finishes its own columns.  The result is
            # validated on the device; the predicated one-sided calls behind it redo the work with running maxima
            # when the status flag was raised (they exit at once otherwise -- no host sync).
            diag, ref, status = be.pair_ref(xi, all_t, ls, off)
            if W == 1:
                row_lse, u, col_lse = be.pair_lse(xi, all_t, ls, ref, status, need_ls, diag=diag, diag_off=off)
            else:
                row_lse, u, col_part = be.pair_lse(xi, all_t, ls, ref, status, need_ls, col_mode=1, diag=diag, diag_off=off)
                parts = _gather_rows(col_part.unsqueeze(0), W, group)           # [W, B_g + 2]
                col_lse = be.merge_col_sums(parts, W * Bl, off, Bl, status)
                work_i.wait()
            be.row_lse(xi, all_t, ls, off, False, need_ls, run_if=status, out_lse=row_lse, out_rowdot=u)
            be.row_lse(xt, all_i, ls, off, False, False, run_if=status, out_lse=col_lse)
            ctx.uv = (u, None) if need_ls else None   # v comes out of the text-side backward kernel
        else:
            r1 = be.row_lse(xi, all_t, ls, off, True, need_ls)       # rows R of S
            if work_i is not None:
                work_i.wait()
            r2 = be.row_lse(xt, all_i, ls, off, False, need_ls)      # columns R of S
            row_lse, diag, col_lse = r1[0], r1[1], r2[0]
            ctx.uv = (r1[2], r2[2]) if need_ls else None
        loss = be.loss_finalize(row_lse, col_lse, diag, ls)      # the rank's local loss
        if W > 1 and not local_loss:
            # reference: one global [B_g, B_g] problem on every rank == mean of the equal-sized rank losses
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
            loss = loss / W

        own_terms_only = W > 1 and local_loss and not gather_with_grad
        ctx.stats_work = None
        if W > 1 and not own_terms_only:
            # the LSE vectors of the other ranks are only needed by backward: start the gather now, wait there
            stats, ctx.stats_work = _gather_rows_async(torch.stack((row_lse, col_lse)).unsqueeze(0), W, group)  # [W, 2, B_l]
        else:
            stats = None   # W == 1, or own-row terms only: the local vectors are all backward needs

        # `stats` is written by the in-flight collective, so it is kept off autograd's version tracking
        ctx.stats = stats
        ctx.save_for_backward(xi, xt, all_i, all_t, ls, row_lse, col_lse, diag)
        ctx.cfg = (bool(local_loss), bool(gather_with_grad), W, off, own_terms_only, group)
        ctx.in_dtypes = (image_features.dtype, text_features.dtype)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape, logit_scale.device) if torch.is_tensor(logit_scale) else None
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        be = _cabi.get_backend()
        xi, xt, all_i, all_t, ls, row_lse, col_lse, diag = ctx.saved_tensors
        stats = ctx.stats
        local_loss, gather_with_grad, W, off, own_terms_only, group = ctx.cfg
        if ctx.stats_work is not None:
            ctx.stats_work.wait()
            ctx.stats_work = None
        if stats is not None:
            row_lse_all = stats[:, 0, :].reshape(-1)
            col_lse_all = stats[:, 1, :].reshape(-1)
        else:
            row_lse_all, col_lse_all = row_lse, col_lse
        Bl = xi.shape[0]
        Bg = W * Bl
        go = grad_out.detach().to(device=xi.device, dtype=torch.float32).reshape(1).contiguous()

        # 1/(2n) of the feature gradients: the true gradient for W=1 and (False, False),
        # W x that otherwise.
        n_feat = Bg if (W == 1 or (not local_loss and not gather_with_grad)) else Bl
        inv_2n = 1.0 / (2.0 * n_feat)
        if own_terms_only:
            w_row, w_col, w_diag = 1.0, 0.0, 1.0
            lse_y_i = lse_y_t = None
        else:
            w_row, w_col, w_diag = 1.0, 1.0, 2.0
            lse_y_i, lse_y_t = col_lse_all, row_lse_all

        need_i, need_t, need_ls = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_ls = need_ls and ctx.ls_meta is not None and ctx.uv is not None
        # after a two-sided forward the text-side softmax-weighted dots v are still missing: the text-side
        # backward kernel emits them as its `rowdot`
        want_v = need_ls and ctx.uv[1] is None
        d_img = d_txt = d_ls = v_bwd = None
        if need_i:
            d_img, _ = be.block_grad(xi, all_t, ls, go, row_lse, lse_y_i, off, w_row, w_col, w_diag, inv_2n, False)
        if need_t:
            d_txt, v_bwd = be.block_grad(xt, all_i, ls, go, col_lse, lse_y_t, off, w_row, w_col, w_diag, inv_2n, want_v)
        elif want_v:
            v_bwd = be.row_lse(xt, all_i, ls, off, False, True)[2]
        if need_ls:
            u, v = ctx.uv
            if v is None:
                v = v_bwd
            n_ls = Bl if (W > 1 and local_loss) else Bg
            t, d_ls = be.dls_finalize(u, v, diag, go, 1.0 / (2.0 * n_ls))
            if W > 1 and not local_loss:
                t = t.clone()
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                d_ls = go[0] * t / (2.0 * n_ls)
            dt, shape, dev = ctx.ls_meta
            d_ls = d_ls.reshape(shape).to(device=dev, dtype=dt)
        else:
            d_ls = None
        if d_img is not None:
            d_img = d_img.to(ctx.in_dtypes[0]) if need_i else None
        if d_txt is not None:
            d_txt = d_txt.to(ctx.in_dtypes[1]) if need_t else None
        return d_img, d_txt, d_ls, None, None, None, None, None


def clip_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale, local_loss: bool = False,
              gather_with_grad: bool = False, rank: int = 0, world_size: int = 1, group=None) -> torch.Tensor:
    """Functional form with the argument validation the reference leaves to torch errors / hangs."""
    if image_features.dim() != 2 or text_features.dim() != 2:
        raise ValueError(f"features must be [B, D]; got {tuple(image_features.shape)} and {tuple(text_features.shape)}")
    if image_features.shape != text_features.shape:
        raise ValueError(f"image/text feature shapes differ: {tuple(image_features.shape)} vs {tuple(text_features.shape)}")
    if image_features.device != text_features.device:
        raise ValueError("image_features and text_features are on different devices")
    if image_features.shape[0] == 0 or image_features.shape[1] == 0:
        raise ValueError("empty batch")
    if torch.is_tensor(logit_scale) and logit_scale.numel() != 1:
        raise ValueError("logit_scale must be a scalar")
    world_size = int(world_size)
    if world_size > 1:
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("world_size > 1 needs an initialised torch.distributed process group")
        pg = dist.get_world_size(group)
        if pg != world_size:
            raise ValueError(f"world_size={world_size} does not match the process group size {pg}")
        if not (0 <= int(rank) < world_size):
            raise ValueError(f"rank {rank} outside [0, {world_size})")
    return ClipLossFunction.apply(image_features, text_features, logit_scale, bool(local_loss),
                                  bool(gather_with_grad), int(rank), world_size, group)
