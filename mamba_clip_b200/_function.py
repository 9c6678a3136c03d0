"""Autograd Function + collective plumbing behind `ClipLoss.forward`.

The reference builds this with implicit autograd over `logit_scale * a @ b.T` and
`F.cross_entropy` (reference src/mamba_clip/loss.py:89-113,142-145) and torch's `_AllGather`.
Here (see DESIGN.md for the decomposition):

  rank r owns rows R of S = ls * I T^T (block "I_r x all T") and columns R of S (block "T_r x all I").
  forward : bf16/f16, D <= 512: ONE two-sided pass over the row block gives row_lse[R] and the sums of every column over
            the rank's rows; at W > 1 one [B_g + 2 + B_l] statistics message per rank is all-gathered and every rank
            finishes all columns.  A device-side status flag chains the robust one-sided kernels behind it (predicated,
            no host sync).  fp32 / D > 512: two one-sided row-LSE calls (rows R, columns R) + an all-gather of the LSEs.
  exchange: features are all-gathered once into contiguous buffers (direct NCCL calls on the compute stream, the image
            gather on a side stream; torch.distributed collectives as the fallback).
  backward: two recompute launches; dI_r and dT_r are each complete on the owning rank -> no gradient collective at
            all; only `local_loss=False` needs one scalar all-reduce for d(logit_scale).
  launch  : eager, or -- opt-in -- CUDA-graph replay of the kernel segments between the collectives.
Per-rank values reproduce the reference exactly, including the W x factors.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch.autograd.function import once_differentiable

from . import _cabi, _nccl

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


class _EagerRunner:
    """Runs every segment directly and allocates fresh buffers (the default path)."""

    def seg(self, name, fn):
        return fn()

    def buffer(self, name, shape, dtype, device):
        return torch.empty(shape, dtype=dtype, device=device)


_EAGER = _EagerRunner()


class _Done:
    """Work handle of a collective that was issued in order on the compute stream: nothing to wait for."""

    @staticmethod
    def wait():
        return None


_comm_streams = {}
_NCCL_DEBUG_TIMEOUT_S = float(os.environ.get("MCLIP_NCCL_TIMEOUT_S", "0") or 0)      # read once at import


class _EventWait:
    """Work handle of a collective issued on the per-device communication stream: `wait()` makes the current stream wait
    for the event recorded right behind it."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


def _on_comm_stream(dev, fn):
    """Run `fn()` (direct NCCL calls) on the device's communication stream, ordered after everything already queued on the
    current stream; returns a handle whose `wait()` makes the current stream wait for it.  EVERY direct NCCL call of the
    loss goes through here: one communicator, one stream, program order -- collectives never run concurrently with each
    other, only next to the compute kernels."""
    cs = _comm_streams.get(dev.index)
    if cs is None:
        cs = _comm_streams[dev.index] = torch.cuda.Stream(device=dev)
    cs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cs):
        fn()
        ev = torch.cuda.Event()
        ev.record(cs)
    if _NCCL_DEBUG_TIMEOUT_S > 0:
        # opt-in debugging aid (MCLIP_NCCL_TIMEOUT_S): the direct communicator has no watchdog, so a rank that never reaches the
        # collective would otherwise show up as a silent hang.  Host-synchronous: not for production runs.
        import time
        t0 = time.monotonic()
        while not ev.query():
            if time.monotonic() - t0 > _NCCL_DEBUG_TIMEOUT_S:
                raise RuntimeError(f"mamba_clip_b200: a direct NCCL collective did not complete within {_NCCL_DEBUG_TIMEOUT_S:.0f} s "
                                   "(a rank is missing from the collective, or ranks disagree on its size)")
            time.sleep(0.0005)
    return _EventWait(ev)


_prep_streams = {}
# MCLIP_F16_PRECOPY = 1 / 0 forces it on / off (read once at import); default: on from 8 ranks.  Measured at C3: on 2 GPUs
# hiding the two 11 us copies behind the forward kernel does NOT pay -- whatever they overlap with on the memory system (the
# image all-gather: 56 -> 113 us; the statistics all-gather: 9 -> 54 us) is an NCCL LL kernel on the critical path of the next
# segment, step 2.07 -> 2.08-2.16 ms; on 8 GPUs (4 MB shards, 0.75 ms steps) it does: 0.755 -> 0.727 ms (tools/r2_n8_matrix.sh).
_F16_PRECOPY = {"1": True, "0": False}.get(os.environ.get("MCLIP_F16_PRECOPY", ""), None)


def _f16_precopy(W: int) -> bool:
    return _F16_PRECOPY if _F16_PRECOPY is not None else W >= 8


def _start_f16_copies(be, run, all_i, all_t, work_i, work_t):
    """bf16 inputs: the backward's dX MMA needs an f16 copy of each `[N, D]` operand (G is f16 * 2^12).  Make both copies
    on a per-device side stream as soon as the (gathered) features exist -- next to the forward kernel, which is tensor
    bound and leaves the memory system idle -- instead of in front of each recompute launch.  -> (y16_i, y16_t, event)."""
    dev = all_t.device
    ps = _prep_streams.get(dev.index)
    if ps is None:
        ps = _prep_streams[dev.index] = torch.cuda.Stream(device=dev)
    ps.wait_stream(torch.cuda.current_stream(dev))
    y16_t = run.buffer("y16_t", tuple(all_t.shape), torch.float16, dev)
    y16_i = run.buffer("y16_i", tuple(all_i.shape), torch.float16, dev)
    with torch.cuda.stream(ps):
        # both copies start once BOTH gathers are done: a bandwidth kernel next to an NCCL LL all-gather slows the gather
        # (measured: the image gather went 56 -> 113 us and delayed the statistics gather queued behind it)
        if work_t is not None:
            work_t.wait()
        if work_i is not None:
            work_i.wait()
        be.to_f16(all_t, out=y16_t)
        be.to_f16(all_i, out=y16_i)
        ev = torch.cuda.Event()
        ev.record(ps)
    return y16_i, y16_t, ev


def _gather_into(out, x, group, comm=None):
    """All-gather into a caller-provided buffer: directly through NCCL on the current stream when a direct
    communicator exists (see _nccl.py), else through torch.distributed.  Returns a handle with `.wait()`."""
    if comm is not None:
        return _on_comm_stream(x.device, lambda: comm.all_gather(out, x))
    try:
        return dist.all_gather_into_tensor(out, x, group=group, async_op=True)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        return dist.all_gather(list(out.chunk(dist.get_world_size(group), dim=0)), x, group=group, async_op=True)


def _small_forward_impl(be, xi, xt, ls, local_loss, rank, W, group, run, send=None):
    """Latency path (B_g <= 1024): [pack + ONE all-gather of (image shard; text shard)] + ONE forward kernel.  Every rank
    evaluates the whole B_g x B_g problem, like the reference does for local_loss=False (loss.py:104-108): no statistics
    exchange and no scalar all-reduce in either direction."""
    Bl, D = xi.shape
    Bg = W * Bl
    dev = xi.device
    if W > 1:
        if send is None:       # (a producer that wrote xi / xt into one [2, Bl, D] buffer passes it: no pack kernel)
            send = be.small_pack(xi, xt, xi.dtype)                     # [2, Bl, D]
        recv = run.buffer("small_recv", (W, 2, Bl, D), xi.dtype, dev)
        comm = _nccl.direct_comm(group, dev) if dev.type == "cuda" else None
        if comm is not None:
            # pack -> gather -> forward kernel are strictly dependent: issue the gather in order on the compute stream
            # (no stream hop, no events).  The general path's communication stream always waits for the compute stream
            # before its own collectives, so calls of the one communicator stay totally ordered across the two paths.
            comm.all_gather(recv.view(W, 2 * Bl * D), send.view(1, 2 * Bl * D))
        else:
            _gather_into(recv.view(W, 2 * Bl * D), send.view(1, 2 * Bl * D), group, None).wait()
        A, Bm, stride, off = recv, recv[0, 1], 2 * Bl * D, int(rank) * Bl
    else:
        A, Bm, stride, off = xi, xt, Bl * D, 0
    lo, hi = (off, off + Bl) if (W > 1 and local_loss) else (0, Bg)
    stats = run.seg("small_fwd", lambda: be.small_forward(A, Bm, Bl, Bg, D, stride, ls, lo, hi))
    return dict(small=True, loss=stats[5 * Bg:5 * Bg + 1].reshape(()), stats=stats, A=A, Bm=Bm, stride=stride, off=off, ls=ls, xi=xi, xt=xt)


def _small_backward_impl(be, st, go, local_loss, gather_with_grad, W, need_ls, run):
    Bl, D = st["xi"].shape
    Bg = W * Bl
    own_terms_only = W > 1 and local_loss and not gather_with_grad
    w = (1.0, 0.0, 1.0) if own_terms_only else (1.0, 1.0, 2.0)
    n_feat = Bg if (W == 1 or (not local_loss and not gather_with_grad)) else Bl
    n_ls = Bl if (W > 1 and local_loss) else Bg
    d_img, d_txt, d_ls = run.seg("small_bwd", lambda: be.small_backward(st["A"], st["Bm"], Bl, Bg, D, st["stride"], st["ls"], go,
                                                                           st["stats"], st["off"], w[0], w[1], w[2],
                                                                           1.0 / (2.0 * n_feat), 1.0 / (2.0 * n_ls)))
    return d_img, d_txt, (d_ls[0] if need_ls else None)


def _forward_impl(be, xi, xt, ls, local_loss, gather_with_grad, rank, W, group, need_ls, run=_EAGER, gathered=None,
                  allow_small=False, for_backward=False):
    if allow_small and gathered is None:      # the caller has already asked be.small_supported(...)
        return _small_forward_impl(be, xi, xt, ls, local_loss, rank, W, group, run)
    """Everything `ClipLoss.forward` launches, on already-cast contiguous inputs.  Returns the state dict that
    `_backward_impl` consumes.  Kernel launches are grouped into segments (`run.seg`) separated by the collectives:
    a segment is pure stream work without host synchronisation, so the graph runner can capture and replay it, while
    the collectives always go through torch.distributed directly."""
    Bl, D = xi.shape
    dev = xi.device
    comm = _nccl.direct_comm(group, dev) if (W > 1 and dev.type == "cuda") else None
    if W > 1:
        if gathered is not None:
            # the producer epilogue already wrote this rank's shards into its slots of the gather buffers (xi / xt are
            # views of them): the all-gathers run in place, no staging copy of the shard
            all_i, all_t = gathered
        else:
            all_t = run.buffer("all_t", (W * Bl, D), xt.dtype, dev)
            all_i = run.buffer("all_i", (W * Bl, D), xi.dtype, dev)
        if comm is not None:
            # ONE communicator, ONE communication stream, program order: the text gather, then the image gather.  The
            # forward kernels only wait for the first; the second runs next to them (never two collectives at once)
            work_t = _on_comm_stream(dev, lambda: comm.all_gather(all_t, xt))
            work_i = _on_comm_stream(dev, lambda: comm.all_gather(all_i, xi))
        else:
            # c10d: both gathers in flight on its own stream; the image gather overlaps the first kernels
            # (a c10d all-gather must not alias its input with a slot of its output: copy the pre-placed shards out)
            work_t = _gather_into(all_t, xt.clone() if gathered is not None else xt, group, None)
            work_i = _gather_into(all_i, xi.clone() if gathered is not None else xi, group, None)
        off = int(rank) * Bl
    else:
        all_i, all_t, off = xi, xt, 0
        work_i = work_t = None

    f16_copies = None
    if (_f16_precopy(W) and for_backward and xi.dtype == torch.bfloat16 and dev.type == "cuda" and hasattr(be, "to_f16") and xi.shape[1] % 8 == 0 and xi.shape[1] <= 768
            and not torch.cuda.is_current_stream_capturing()
            and not (W == 1 and getattr(be, "fused_supported", lambda *_: False)(xi, all_t))):
        # (c10d handles -- the fallback collectives -- are waited for on the compute stream only)
        f16_copies = _start_f16_copies(be, run, all_i, all_t, work_i if isinstance(work_i, _EventWait) else None,
                                       work_t if isinstance(work_t, _EventWait) else None) \
            if (W == 1 or isinstance(work_t, _EventWait)) else None

    # u, v: softmax-weighted raw dots (sum_j P_ij <x_i, y_j>) of the two blocks -- all d(logit_scale) needs
    row_lse_all = col_lse_all = None      # full-length LSE vectors, when the forward already leaves them on every rank
    if getattr(be, "pair_supported", lambda *_: False)(xi, all_t):
        # two-sided forward: one pass over the rank's row block S_r = ls * I_r T_all^T gives its row LSEs and the
        # sums of every column over its rows (1 GEMM unit instead of 2).  At W > 1 the per-rank column sums are
        # all-gathered (4 * (B_g + 2) bytes per rank) and each rank finishes its own columns.  The result is
        # validated on the device; the predicated one-sided calls behind it redo the work with running maxima
        # when the status flag was raised (they exit at once otherwise -- no host sync).
        Bg = W * Bl
        if W == 1:
            def seg_a():
                diag, ref, status = be.pair_ref(xi, all_t, ls, off)
                row_lse, u, col_lse = be.pair_lse(xi, all_t, ls, ref, status, need_ls, diag=diag, diag_off=off)
                be.row_lse(xi, all_t, ls, off, False, need_ls, run_if=status, out_lse=row_lse, out_rowdot=u)
                be.row_lse(xt, all_i, ls, off, False, False, run_if=status, out_lse=col_lse)
                return diag, row_lse, u, col_lse, be.loss_finalize(row_lse, col_lse, diag, ls)
            diag, row_lse, u, col_lse, loss = run.seg("fwd_a", seg_a)
        else:
            # ONE message per rank: [column sums over its rows (B_g), reference, status, its row LSEs (B_l)].  After the
            # all-gather every rank holds everything: it finishes ALL columns itself and has every rank's row LSEs, so
            # backward needs no further exchange.  The status words are ORed on every rank alike; if any is set, every
            # rank redoes the full row and column statistics with the robust one-sided kernels (W x redundant, rare).
            # the positive pairs of the rank's rows are its own text rows: the exponent reference does not need the gather
            diag, ref, status = run.seg("fwd_ref", lambda: be.pair_ref(xi, xt, ls, 0))
            work_t.wait()

            def seg_a():
                msg = torch.empty(Bg + 2 + Bl, dtype=torch.float32, device=dev)
                u_all = torch.empty(Bg, dtype=torch.float32, device=dev) if need_ls else None   # only [off, off + B_l) is read
                be.pair_lse(xi, all_t, ls, ref, status, need_ls, col_mode=1, diag=diag, diag_off=off, out_msg=msg,
                            out_rowdot=u_all[off:off + Bl] if need_ls else None)
                return msg, u_all
            msg, u_all = run.seg("fwd_a", seg_a)
            parts = run.buffer("col_parts", (W, Bg + 2 + Bl), torch.float32, dev)
            _gather_into(parts, msg.unsqueeze(0), group, comm).wait()
            work_i.wait()

            def seg_b():
                col_all = be.merge_col_sums(parts, Bg, 0, Bg, status)
                row_all = parts[:, Bg + 2:].reshape(-1)                      # [W, B_l] strided -> contiguous [B_g]
                be.row_lse(all_i, all_t, ls, 0, False, need_ls, run_if=status, out_lse=row_all, out_rowdot=u_all)
                be.row_lse(all_t, all_i, ls, 0, False, False, run_if=status, out_lse=col_all)
                return row_all, col_all, be.loss_finalize(row_all[off:off + Bl], col_all[off:off + Bl], diag, ls)
            row_lse_all, col_lse_all, loss = run.seg("fwd_b", seg_b)
            row_lse, col_lse = row_lse_all[off:off + Bl], col_lse_all[off:off + Bl]
            u = u_all[off:off + Bl] if need_ls else None
        uv = (u, None) if need_ls else None   # v comes out of the text-side backward kernel
    else:
        if work_t is not None:
            work_t.wait()
        if work_i is not None:
            work_i.wait()

        def seg_a():
            r1 = be.row_lse(xi, all_t, ls, off, True, need_ls)       # rows R of S
            r2 = be.row_lse(xt, all_i, ls, off, False, need_ls)      # columns R of S
            return r1, r2, be.loss_finalize(r1[0], r2[0], r1[1], ls)
        r1, r2, loss = run.seg("fwd_a", seg_a)
        row_lse, diag, col_lse = r1[0], r1[1], r2[0]
        uv = (r1[2], r2[2]) if need_ls else None
    if W > 1 and not local_loss:
        # reference: one global [B_g, B_g] problem on every rank == mean of the equal-sized rank losses
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        loss = loss / W

    own_terms_only = W > 1 and local_loss and not gather_with_grad
    stats = stats_work = None
    if W > 1 and not own_terms_only and row_lse_all is None:
        # the LSE vectors of the other ranks are only needed by backward: start the gather now, wait there
        stats = run.buffer("stats", (W, 2, Bl), torch.float32, dev)
        stats_work = _gather_into(stats, torch.stack((row_lse, col_lse)).unsqueeze(0), group, comm)
    if f16_copies is not None:
        # join: from here on the copies are ordered before everything the compute stream does next, so the buffers (allocated on
        # the compute stream) need no record_stream bookkeeping even if the loss is dropped without a backward
        torch.cuda.current_stream(dev).wait_event(f16_copies[2])
    return dict(loss=loss, xi=xi, xt=xt, all_i=all_i, all_t=all_t, ls=ls, row_lse=row_lse, col_lse=col_lse, diag=diag,
                uv=uv, stats=stats, stats_work=stats_work, off=off, own_terms_only=own_terms_only,
                row_lse_all=row_lse_all, col_lse_all=col_lse_all, f16_copies=f16_copies)


def _backward_impl(be, st, go, local_loss, gather_with_grad, W, group, need_i, need_t, need_ls, run=_EAGER, rows=None):
    """Everything `ClipLoss` launches in backward.  -> (d_image, d_text, d_logit_scale as a 0-dim f32 tensor).
    `rows = (lo, hi)`: only the local rows [lo, hi) need feature gradients (gradient accumulation: the other rows are
    cached, detached features of earlier micro-batches) -- the two recompute launches shrink to that row range."""
    if st.get("small"):
        d_img, d_txt, d_ls = _small_backward_impl(be, st, go, local_loss, gather_with_grad, W, need_ls, run)
        return (d_img if need_i else None), (d_txt if need_t else None), d_ls
    xi, xt, all_i, all_t, ls = st["xi"], st["xt"], st["all_i"], st["all_t"], st["ls"]
    row_lse, col_lse, diag, off = st["row_lse"], st["col_lse"], st["diag"], st["off"]
    own_terms_only = st["own_terms_only"]
    if st["stats_work"] is not None:
        st["stats_work"].wait()
        st["stats_work"] = None
    stats = st["stats"]
    Bl = xi.shape[0]
    Bg = W * Bl
    y16_i = y16_t = None
    if st.get("f16_copies") is not None:
        y16_i, y16_t, _ = st["f16_copies"]          # (the forward already joined the side stream)

    # 1/(2n) of the feature gradients: the true gradient for W=1 and (False, False),
    # W x that otherwise.
    n_feat = Bg if (W == 1 or (not local_loss and not gather_with_grad)) else Bl
    inv_2n = 1.0 / (2.0 * n_feat)
    need_ls = need_ls and st["uv"] is not None
    # after a two-sided forward the text-side softmax-weighted dots v are still missing: the text-side
    # backward kernel emits them as its `rowdot`
    want_v = need_ls and st["uv"][1] is None
    n_ls = Bl if (W > 1 and local_loss) else Bg

    def seg():
        if st.get("row_lse_all") is not None:        # two-sided forward at W > 1: every rank already holds both vectors
            row_lse_all, col_lse_all = st["row_lse_all"], st["col_lse_all"]
        elif stats is not None:
            row_lse_all = stats[:, 0, :].reshape(-1)
            col_lse_all = stats[:, 1, :].reshape(-1)
        else:
            row_lse_all, col_lse_all = row_lse, col_lse
        if own_terms_only:
            w_row, w_col, w_diag = 1.0, 0.0, 1.0
            lse_y_i = lse_y_t = None
        else:
            w_row, w_col, w_diag = 1.0, 1.0, 2.0
            lse_y_i, lse_y_t = col_lse_all, row_lse_all
        d_img = d_txt = t = d_ls = v_bwd = None
        kw_t = {"y16": y16_t} if y16_t is not None else {}      # f16 copies made next to the forward kernel
        kw_i = {"y16": y16_i} if y16_i is not None else {}
        if (W == 1 and rows is None and need_i and need_t and not own_terms_only
                and getattr(be, "fused_supported", lambda *_: False)(xi, all_t)):
            # shared-recompute backward: one recompute of S feeds dI = G T and dT = G^T I (4 GEMM units per step
            # instead of 5); d(logit_scale) from Euler's identity sum_i <I_i, dI_i> = ls * d ls (xdot, f32 accumulators)
            d_img, d_txt, xdot = be.fused_grad(xi, all_t, ls, go, row_lse, col_lse, off, inv_2n)
            if need_ls:
                t, d_ls = be.dls_finalize(xdot, None, None, go, 1.0 / (2.0 * n_ls))
            return d_img, d_txt, t, d_ls
        if rows is not None:
            lo, hi = rows
            if need_i:
                d_img, _ = be.block_grad(xi[lo:hi], all_t, ls, go, row_lse[lo:hi], lse_y_i, off + lo, w_row, w_col, w_diag,
                                         inv_2n, False, **kw_t)
            if need_t:
                d_txt, _ = be.block_grad(xt[lo:hi], all_i, ls, go, col_lse[lo:hi], lse_y_t, off + lo, w_row, w_col, w_diag,
                                         inv_2n, False, **kw_i)
            if want_v:      # v of ALL local rows feeds d(logit_scale): one forward-only pass of the text side
                v_bwd = be.row_lse(xt, all_i, ls, off, False, True)[2]
        else:
            if need_i:
                d_img, _ = be.block_grad(xi, all_t, ls, go, row_lse, lse_y_i, off, w_row, w_col, w_diag, inv_2n, False, **kw_t)
            if need_t:
                d_txt, v_bwd = be.block_grad(xt, all_i, ls, go, col_lse, lse_y_t, off, w_row, w_col, w_diag, inv_2n, want_v, **kw_i)
            elif want_v:
                v_bwd = be.row_lse(xt, all_i, ls, off, False, True)[2]
        if need_ls:
            u, v = st["uv"]
            if v is None:
                v = v_bwd
            t, d_ls = be.dls_finalize(u, v, diag, go, 1.0 / (2.0 * n_ls))
        return d_img, d_txt, t, d_ls
    d_img, d_txt, t, d_ls = run.seg("bwd", seg)
    if need_ls and W > 1 and not local_loss:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        d_ls = go[0] * t / (2.0 * n_ls)
    return d_img, d_txt, d_ls


# ----------------------------------------------------------------------------------------------------
# CUDA-graph replay of the launch segments (opt-in)
# ----------------------------------------------------------------------------------------------------
# One loss step is ~20 kernel launches and up to 4 collectives.  At the headline shapes on 8 GPUs the kernels of a rank
# take ~0.5 ms, which is less than the Python / ctypes time needed to enqueue them: the step is host-bound.  With
# graphs enabled, the kernel segments of a given (shape, dtype, mode) are captured once (after two eager warm-up calls)
# and replayed; per call the host copies the inputs into the static buffers, issues the collectives (always eagerly,
# between the segments, into static buffers: no NCCL kernel is ever captured) and replays.
_GRAPHS_ENABLED = os.environ.get("MCLIP_CUDA_GRAPHS", "0") == "1"
_GRAPH_WARMUP = 2
# Replay only pays where the step is launch-bound.  From ~2^37 multiply-adds per rank and pass (B_l * B_g * D; the
# headline C3 shape is 2^36 at 8 GPUs, 2^37 at 4 and 2^39 at one) the kernels outlast the host by a wide margin, and
# the static input copies / output clones of the graph path (4 x B_l x D elements) would only add traffic.
_GRAPH_MAX_WORK = int(os.environ.get("MCLIP_GRAPH_MAX_WORK", str(1 << 37)))
_graph_cache = {}
_GRAPH_CACHE_MAX = 8


def enable_cuda_graphs(flag: bool = True) -> None:
    """Replay captured CUDA graphs for the kernel segments of `ClipLoss` forward/backward (per shape/dtype/mode, after
    two eager calls).  Same kernels, same results; removes the host launch overhead.  Also: MCLIP_CUDA_GRAPHS=1."""
    global _GRAPHS_ENABLED
    _GRAPHS_ENABLED = bool(flag)
    if not flag:
        _graph_cache.clear()


def cuda_graphs_enabled() -> bool:
    return _GRAPHS_ENABLED


_SAVED_KEYS = ("xi", "xt", "all_i", "all_t", "ls", "row_lse", "col_lse", "diag")


class _PendingToken:
    """Lives on the autograd ctx of a graphed forward: if the ctx dies without a backward (loss discarded), the
    graph's buffers are released for the next forward."""

    def __init__(self, gl, generation):
        self.gl, self.generation = gl, generation

    def __del__(self):
        gl = self.gl
        if gl is not None and gl.pending and gl.generation == self.generation:
            gl.pending = False


class _GraphedLoss:
    """Static buffers + the captured segment graphs for one (device, shape, dtype, mode) key."""

    def __init__(self, xi, xt, cfg, need_ls):
        self.cfg = cfg                       # (local_loss, gather_with_grad, rank, W, group)
        self.need_ls = need_ls
        self.calls = 0
        self.xi = torch.empty_like(xi)
        self.xt = torch.empty_like(xt)
        self.ls = torch.empty(1, dtype=torch.float32, device=xi.device)
        self.go = torch.ones(1, dtype=torch.float32, device=xi.device)
        self.graphs = {}                     # segment name -> (CUDAGraph, outputs)
        self.buffers = {}
        self.state = None
        self.pending = False                 # a forward is waiting for its backward: the statics must not be touched
        self.generation = 0
        self._prefix = ""
        # One graph per direction (collectives included) when every collective of the step is a direct in-stream NCCL
        # call; otherwise one graph per kernel segment with the collectives issued eagerly between them.
        local_loss, _, _, W, group = cfg
        self.whole = W == 1 or (os.environ.get("MCLIP_GRAPH_NCCL", "0") == "1" and local_loss and
                                _nccl.direct_comm(group, xi.device) is not None)

    # runner interface ---------------------------------------------------------------------------
    def seg(self, name, fn):
        key = self._prefix + name
        hit = self.graphs.get(key)
        if hit is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                outs = fn()
            hit = self.graphs[key] = (g, outs)
        hit[0].replay()
        return hit[1]

    def buffer(self, name, shape, dtype, device):
        b = self.buffers.get(name)
        if b is None:
            b = self.buffers[name] = torch.empty(shape, dtype=dtype, device=device)
        return b

    # ---------------------------------------------------------------------------------------------
    def run_forward(self, be, xi, xt, ls):
        local_loss, gwg, rank, W, group = self.cfg
        self.xi.copy_(xi)
        self.xt.copy_(xt)
        self.ls.copy_(ls)
        self._prefix = ""
        if self.whole:
            if "fwd" not in self.graphs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    st = _forward_impl(be, self.xi, self.xt, self.ls, local_loss, gwg, rank, W, group, self.need_ls)
                self.graphs["fwd"] = (g, st)
            g, self.state = self.graphs["fwd"]
            g.replay()
        else:
            self.state = _forward_impl(be, self.xi, self.xt, self.ls, local_loss, gwg, rank, W, group, self.need_ls, run=self,
                                       for_backward=True)
        self.generation += 1
        return self.state["loss"].clone()

    def run_backward(self, be, go, need_i, need_t, need_ls):
        local_loss, gwg, rank, W, group = self.cfg
        self.go.copy_(go)
        self._prefix = f"{int(need_i)}{int(need_t)}{int(need_ls)}:"
        if self.whole:
            key = self._prefix + "bwd_whole"
            if key not in self.graphs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    o = _backward_impl(be, self.state, self.go, local_loss, gwg, W, group, need_i, need_t, need_ls)
                self.graphs[key] = (g, o)
            g, outs = self.graphs[key]
            g.replay()
        else:
            outs = _backward_impl(be, self.state, self.go, local_loss, gwg, W, group, need_i, need_t, need_ls, run=self)
        return tuple(None if o is None else o.clone() for o in outs)


def _as_compute(t: torch.Tensor, cdt: torch.dtype) -> torch.Tensor:
    """Detached, contiguous, in the compute dtype -- without the no-op `.to()` / `.contiguous()` calls when the tensor already
    is (the latency configurations are host-bound: every tensor-method call is ~1-3 us of their step)."""
    t = t.detach()
    if t.dtype != cdt:
        t = t.to(cdt)
    return t if t.is_contiguous() else t.contiguous()


def _as_scalar(x, dev) -> torch.Tensor:
    """[1] f32 device tensor view / copy of a 0-dim tensor or Python number."""
    if torch.is_tensor(x):
        x = x.detach()
        if x.dtype != torch.float32 or x.device != dev:
            x = x.to(device=dev, dtype=torch.float32)
        return x.reshape(1)
    return torch.full((1,), float(x), dtype=torch.float32, device=dev)


class ClipLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, local_loss, gather_with_grad, rank, world_size,
                group):
        be = _cabi.get_backend()
        dev = image_features.device
        cdt = _compute_dtype(image_features)
        xi = _as_compute(image_features, cdt)
        xt = _as_compute(text_features, cdt)
        ls = _as_scalar(logit_scale, dev)
        W = int(world_size)
        need_ls = torch.is_tensor(logit_scale) and logit_scale.requires_grad
        wants_grad = need_ls or image_features.requires_grad or text_features.requires_grad

        ctx.graphed = None
        # the latency path is two kernels: replaying them from a graph would only add the static-buffer copies
        small = getattr(be, "small_supported", lambda *_: False)(xi.shape[0], W * xi.shape[0], xi.shape[1], cdt)
        if (_GRAPHS_ENABLED and not small and dev.type == "cuda" and _cabi._override is None
                and xi.shape[0] * xi.shape[0] * W * xi.shape[1] < _GRAPH_MAX_WORK
                and not torch.cuda.is_current_stream_capturing()):
            key = (dev.index, tuple(xi.shape), cdt, bool(local_loss), bool(gather_with_grad), int(rank), W, id(group), need_ls)
            gl = _graph_cache.get(key)
            if gl is None:
                while len(_graph_cache) >= _GRAPH_CACHE_MAX:       # static buffers + graphs per key: keep it bounded
                    _graph_cache.pop(next(iter(_graph_cache)))
                gl = _graph_cache[key] = _GraphedLoss(xi, xt, (bool(local_loss), bool(gather_with_grad), int(rank), W, group), need_ls)
            gl.calls += 1
            if gl.calls > _GRAPH_WARMUP and not gl.pending:
                loss = gl.run_forward(be, xi, xt, ls)
                if wants_grad:
                    gl.pending = True
                    ctx.graphed = gl
                    ctx.generation = gl.generation
                    ctx.token = _PendingToken(gl, gl.generation)
                ctx.cfg = (bool(local_loss), bool(gather_with_grad), W, group)
                ctx.in_dtypes = (image_features.dtype, text_features.dtype)
                ctx.ls_meta = (logit_scale.dtype, logit_scale.shape, logit_scale.device) if torch.is_tensor(logit_scale) else None
                return loss

        st = _forward_impl(be, xi, xt, ls, bool(local_loss), bool(gather_with_grad), int(rank), W, group, need_ls, allow_small=small,
                           for_backward=wants_grad)
        loss = st.pop("loss")
        if not st.get("small"):
            # tensors go through save_for_backward (in-place modification checks); `stats` is written by an in-flight
            # collective, so it is kept off autograd's version tracking together with the non-tensor state
            ctx.save_for_backward(*(st.pop(k) for k in _SAVED_KEYS))
        ctx.state = st
        ctx.cfg = (bool(local_loss), bool(gather_with_grad), W, group)
        ctx.in_dtypes = (image_features.dtype, text_features.dtype)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape, logit_scale.device) if torch.is_tensor(logit_scale) else None
        return loss

    @staticmethod
    @once_differentiable     # the gradients come from raw kernels: a double backward must raise, not return zeros
    def backward(ctx, grad_out):
        be = _cabi.get_backend()
        local_loss, gather_with_grad, W, group = ctx.cfg
        need_i, need_t, need_ls = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_ls = need_ls and ctx.ls_meta is not None
        gl = ctx.graphed
        if gl is not None:
            if ctx.generation != gl.generation:
                raise RuntimeError("mamba_clip_b200: the CUDA-graph buffers of this loss were overwritten by a later forward; "
                                   "call backward before the next forward of the same shape, or disable CUDA graphs")
            go = grad_out.detach().to(device=gl.go.device, dtype=torch.float32).reshape(1)
            d_img, d_txt, d_ls = gl.run_backward(be, go, need_i, need_t, need_ls)
            gl.pending = False   # (a second backward through retain_graph stays valid until the next forward)
        else:
            st = dict(ctx.state)
            if not st.get("small"):
                st.update(zip(_SAVED_KEYS, ctx.saved_tensors))
            go = _as_scalar(grad_out, st["xi"].device)
            d_img, d_txt, d_ls = _backward_impl(be, st, go, local_loss, gather_with_grad, W, group, need_i, need_t, need_ls)
        if d_ls is not None:
            dt, shape, dev = ctx.ls_meta
            d_ls = d_ls.reshape(shape)
            if d_ls.dtype != dt or d_ls.device != dev:
                d_ls = d_ls.to(device=dev, dtype=dt)
        if d_img is not None and d_img.dtype != ctx.in_dtypes[0]:
            d_img = d_img.to(ctx.in_dtypes[0])
        if d_txt is not None and d_txt.dtype != ctx.in_dtypes[1]:
            d_txt = d_txt.to(ctx.in_dtypes[1])
        return d_img, d_txt, d_ls, None, None, None, None, None


class ClipLossFromProjectionsFunction(torch.autograd.Function):
    """Producer epilogue fused into the gather prologue (reference model.py:1011-1017,1051 +
    loss.py:16-44): raw fp32 tower projections -> ONE kernel per tower that L2-normalises each row, rounds it to the
    16-bit compute dtype and writes it straight into this rank's slot of the all-gather buffer -> in-place NCCL all-gather
    -> the fused loss.  Backward: the loss kernels' feature gradients go through the normalisation backward kernel and
    come out as fp32 gradients of the raw projections.  Same values as `ClipLoss(normalize_features(a), normalize_features(b), ls)`."""

    @staticmethod
    def forward(ctx, image_proj, text_proj, logit_scale, dtype, eps, local_loss, gather_with_grad, rank, world_size, group):
        be = _cabi.get_backend()
        dev = image_proj.device
        pi = image_proj.detach().to(torch.float32).contiguous()
        pt = text_proj.detach().to(torch.float32).contiguous()
        Bl, D = pi.shape
        W = int(world_size)
        off = int(rank) * Bl if W > 1 else 0
        if torch.is_tensor(logit_scale):
            ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        else:
            ls = torch.full((1,), float(logit_scale), dtype=torch.float32, device=dev)
        need_ls = torch.is_tensor(logit_scale) and logit_scale.requires_grad
        if be.small_supported(Bl, W * Bl, D, dtype):
            # latency path: normalise straight into the [image shard; text shard] send buffer of its single all-gather
            send = torch.empty((2, Bl, D), dtype=dtype, device=dev)
            xi = be.normalize_rows(pi, dtype, eps, out=send[0])
            xt = be.normalize_rows(pt, dtype, eps, out=send[1])
            st = _small_forward_impl(be, xi, xt, ls, bool(local_loss), int(rank), W, group, _EAGER, send=send)
        else:
            all_i = torch.empty((W * Bl, D), dtype=dtype, device=dev)
            all_t = torch.empty((W * Bl, D), dtype=dtype, device=dev)
            xi = be.normalize_rows(pi, dtype, eps, out=all_i[off:off + Bl])
            xt = be.normalize_rows(pt, dtype, eps, out=all_t[off:off + Bl])
            st = _forward_impl(be, xi, xt, ls, bool(local_loss), bool(gather_with_grad), int(rank), W, group, need_ls,
                               gathered=(all_i, all_t) if W > 1 else None)
        loss = st.pop("loss")
        if st.get("small"):
            ctx.save_for_backward(pi, pt)
        else:
            ctx.save_for_backward(pi, pt, *(st.pop(k) for k in _SAVED_KEYS))
        ctx.state = st
        ctx.eps = eps
        ctx.cfg = (bool(local_loss), bool(gather_with_grad), W, group)
        ctx.in_dtypes = (image_proj.dtype, text_proj.dtype)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape, logit_scale.device) if torch.is_tensor(logit_scale) else None
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        be = _cabi.get_backend()
        local_loss, gather_with_grad, W, group = ctx.cfg
        need_i, need_t, need_ls = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_ls = need_ls and ctx.ls_meta is not None
        pi, pt, *saved = ctx.saved_tensors
        st = dict(ctx.state)
        if not st.get("small"):
            st.update(zip(_SAVED_KEYS, saved))
        go = grad_out.detach().to(device=pi.device, dtype=torch.float32).reshape(1).contiguous()
        d_img, d_txt, d_ls = _backward_impl(be, st, go, local_loss, gather_with_grad, W, group, need_i, need_t, need_ls)
        if d_img is not None:
            d_img = be.normalize_rows_bwd(pi, d_img.contiguous(), ctx.eps).to(ctx.in_dtypes[0])
        if d_txt is not None:
            d_txt = be.normalize_rows_bwd(pt, d_txt.contiguous(), ctx.eps).to(ctx.in_dtypes[1])
        if d_ls is not None:
            dt, shape, dev = ctx.ls_meta
            d_ls = d_ls.reshape(shape).to(device=dev, dtype=dt)
        return (d_img, d_txt, d_ls) + (None,) * 7


class ClipLossChunkFunction(torch.autograd.Function):
    """Loss over `[cached micro-batches with rows [lo, hi) replaced by the live micro-batch]`, differentiable w.r.t. the
    live micro-batch and `logit_scale` only.  Same forward as `ClipLossFunction` on the concatenation; the backward
    recompute launches cover only the live rows (1/accum_freq of the work)."""

    @staticmethod
    def forward(ctx, image_j, text_j, logit_scale, image_full, text_full, lo, local_loss, gather_with_grad, rank,
                world_size, group):
        be = _cabi.get_backend()
        dev = image_j.device
        cdt = _compute_dtype(image_j)
        hi = lo + image_j.shape[0]
        xi = image_full.detach().to(cdt).contiguous().clone()
        xt = text_full.detach().to(cdt).contiguous().clone()
        xi[lo:hi] = image_j.detach().to(cdt)
        xt[lo:hi] = text_j.detach().to(cdt)
        if torch.is_tensor(logit_scale):
            ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        else:
            ls = torch.full((1,), float(logit_scale), dtype=torch.float32, device=dev)
        W = int(world_size)
        need_ls = torch.is_tensor(logit_scale) and logit_scale.requires_grad
        st = _forward_impl(be, xi, xt, ls, bool(local_loss), bool(gather_with_grad), int(rank), W, group, need_ls)
        loss = st.pop("loss")
        ctx.save_for_backward(*(st.pop(k) for k in _SAVED_KEYS))
        ctx.state = st
        ctx.rows = (lo, hi)
        ctx.cfg = (bool(local_loss), bool(gather_with_grad), W, group)
        ctx.in_dtypes = (image_j.dtype, text_j.dtype)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape, logit_scale.device) if torch.is_tensor(logit_scale) else None
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        be = _cabi.get_backend()
        local_loss, gather_with_grad, W, group = ctx.cfg
        need_i, need_t, need_ls = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_ls = need_ls and ctx.ls_meta is not None
        st = dict(ctx.state)
        st.update(zip(_SAVED_KEYS, ctx.saved_tensors))
        go = grad_out.detach().to(device=st["xi"].device, dtype=torch.float32).reshape(1).contiguous()
        d_img, d_txt, d_ls = _backward_impl(be, st, go, local_loss, gather_with_grad, W, group, need_i, need_t, need_ls,
                                            rows=ctx.rows)
        if d_ls is not None:
            dt, shape, dev = ctx.ls_meta
            d_ls = d_ls.reshape(shape).to(device=dev, dtype=dt)
        if d_img is not None:
            d_img = d_img.to(ctx.in_dtypes[0])
        if d_txt is not None:
            d_txt = d_txt.to(ctx.in_dtypes[1])
        return (d_img, d_txt, d_ls) + (None,) * 8


_checked_shards = set()


def _check_equal_shards(x: torch.Tensor, group) -> None:
    """Every rank must bring the same [B_l, D] and dtype: the flat all-gather and the fixed row offsets `rank * B_l`
    silently corrupt (or hang) otherwise, where the reference would fail inside torch.  Checked with one tiny MIN/MAX
    all-reduce the first time a (group, shape, dtype) is seen -- one host synchronisation per new shape, none per step."""
    key = (group, tuple(x.shape), x.dtype, x.device)      # the group object itself: its id cannot be reused while the key lives
    if key in _checked_shards:
        return
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}.get(x.dtype, 3)
    me = torch.tensor([x.shape[0], x.shape[1], code], dtype=torch.int64)
    dev = x.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    v = torch.stack((me, -me)).to(dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
    mx, mn = v[0].cpu(), -v[1].cpu()
    if not torch.equal(mx, mn):
        raise ValueError(f"every rank must pass features of the same shape and dtype; across the group B_l ranges over "
                         f"[{int(mn[0])}, {int(mx[0])}], D over [{int(mn[1])}, {int(mx[1])}] (this rank: {tuple(x.shape)}, {x.dtype})")
    _checked_shards.add(key)


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
        _check_equal_shards(image_features, group)
    return ClipLossFunction.apply(image_features, text_features, logit_scale, bool(local_loss),
                                  bool(gather_with_grad), int(rank), world_size, group)
