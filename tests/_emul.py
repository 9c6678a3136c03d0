"""Oracle-backed stand-in for the four C-ABI primitives, for CPU tests of the host logic only
(world_size-2/4 gloo).  Lives under tests/: the product never imports it."""
import torch

from oracle import clip_oracle as O


class EmulatedBackend:
    name = "oracle-emulation"

    def __init__(self):
        self.calls = []

    def row_lse(self, X, Y, ls, diag_off, want_diag, want_rowdot=False, run_if=None, out_lse=None, out_rowdot=None):
        if run_if is not None and int(run_if) == 0:
            self.calls.append("row_lse(skipped)")      # predicated off: nothing is written
            return (out_lse, None, out_rowdot) if want_rowdot else (out_lse, None)
        self.calls.append("row_lse")
        lse, diag = O.block_row_lse(X, Y, float(ls), diag_off)
        out = (lse.float(), (diag.float() if want_diag else None))
        if want_rowdot:
            C = X.double() @ Y.double().T
            P = torch.exp(float(ls) * C - lse[:, None])
            out = out + ((P * C).sum(dim=1).float(),)
        if run_if is not None:                          # predicated on: results land in the caller's buffers
            out_lse.copy_(out[0])
            if want_rowdot:
                out_rowdot.copy_(out[2])
        return out

    def block_grad(self, X, Y, ls, go, lse_x, lse_y, diag_off, w_row, w_col, w_diag, inv_2n, want_rowdot=True):
        self.calls.append("block_grad")
        alpha = float(go) * float(ls) * inv_2n
        dX, rowdot = O.block_grad(X, Y, float(ls), lse_x, lse_y, diag_off, w_row, w_col, w_diag, alpha)
        return dX.to(X.dtype), rowdot.float()

    def loss_finalize(self, row_lse, col_lse, diag, ls):
        self.calls.append("loss_finalize")
        n = row_lse.numel()
        return ((row_lse.double() + col_lse.double() - 2 * float(ls) * diag.double()).sum() / (2 * n)).float()

    def dls_finalize(self, u, v, diag, go, scale):
        self.calls.append("dls_finalize")
        t = u.double().sum()
        if v is not None:
            t = t + v.double().sum()
        if diag is not None:
            t = t - 2 * diag.double().sum()
        return t.float(), (float(go) * scale * t).float()

    def launch_count(self):
        return len(self.calls)


class EmulatedFusedBackend(EmulatedBackend):
    """Adds the oracle statement of mclip_fused_grad (include/mclip_b200.h): both gradients of the full-weight block and
    xdot[i] = sum_j G_ij <x_i, y_j>, so that the W = 1 shared-recompute branch of the host logic runs on CPU."""

    def fused_supported(self, X, Y):
        return True

    def fused_grad(self, X, Y, ls, go, lse_x, lse_y, diag_off, inv_2n):
        self.calls.append("fused_grad")
        alpha = float(go) * float(ls) * inv_2n
        Xd, Yd = X.double(), Y.double()
        C = Xd @ Yd.T
        S = float(ls) * C
        G = torch.exp(S - lse_x.double()[:, None]) + torch.exp(S - lse_y.double()[None, :])
        i = torch.arange(X.shape[0])
        j = i + diag_off
        ok = (j >= 0) & (j < Y.shape[0])
        G[i[ok], j[ok]] -= 2.0
        return (alpha * (G @ Yd)).to(X.dtype), (alpha * (G.T @ Xd)).to(X.dtype), (G * C).sum(dim=1).float()


class EmulatedSmallBackend(EmulatedBackend):
    """Adds the oracle statement of the latency-path primitives (include/mclip_b200.h: mclip_small_pack / _forward /
    _backward) so that its host logic -- ONE packed all-gather, the blocked [W][2][Bl][D] layout, the per-mode row ranges,
    weights and scale factors, no scalar collectives -- runs on CPU under gloo."""

    def small_supported(self, Bl, Bg, D, dtype):
        return Bg <= 1024 and D <= 512

    def small_pack(self, a, b, out_dtype):
        self.calls.append("small_pack")
        return torch.stack((a, b)).to(out_dtype)

    @staticmethod
    def _rows(A, Bm, Bl, Bg, D, stride):
        W = Bg // Bl
        if W == 1:
            return A.reshape(Bg, D).double(), Bm.reshape(Bg, D).double()
        recv = A.reshape(W, 2, Bl, D)
        return recv[:, 0].reshape(Bg, D).double(), recv[:, 1].reshape(Bg, D).double()

    def small_forward(self, A, Bm, Bl, Bg, D, blk_stride, ls, lo, hi):
        self.calls.append("small_forward")
        I, T = self._rows(A, Bm, Bl, Bg, D, blk_stride)
        st = O.global_stats(I, T, float(ls))
        stats = torch.empty(5 * Bg + 2, dtype=torch.float32)
        for k, v in enumerate((st.row_lse, st.col_lse, st.diag, st.u, st.v)):
            stats[k * Bg:(k + 1) * Bg] = v.float()
        per_row = (st.row_lse - float(ls) * st.diag) + (st.col_lse - float(ls) * st.diag)
        stats[5 * Bg] = float(per_row[lo:hi].sum() / (2 * (hi - lo)))
        stats[5 * Bg + 1] = float((st.u + st.v - 2 * st.diag)[lo:hi].sum())
        return stats

    def small_backward(self, A, Bm, Bl, Bg, D, blk_stride, ls, go, stats, off, w_row, w_col, w_diag, inv_2n, dls_scale):
        self.calls.append("small_backward")
        I, T = self._rows(A, Bm, Bl, Bg, D, blk_stride)
        row_lse, col_lse = stats[:Bg].double(), stats[Bg:2 * Bg].double()
        alpha = float(go) * float(ls) * inv_2n
        dA, _ = O.block_grad(I[off:off + Bl], T, float(ls), row_lse[off:off + Bl], col_lse, off, w_row, w_col, w_diag, alpha)
        dB, _ = O.block_grad(T[off:off + Bl], I, float(ls), col_lse[off:off + Bl], row_lse, off, w_row, w_col, w_diag, alpha)
        dls = torch.tensor([float(go) * dls_scale * float(stats[5 * Bg + 1])])
        return dA.to(A.dtype), dB.to(A.dtype), dls


LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453
SUM_LO, SUM_HI = 2.0 ** -75, 2.0 ** 120


class EmulatedPairBackend(EmulatedBackend):
    """Adds the oracle statement of the two-sided forward primitives (include/mclip_b200.h: mclip_pair_ref,
    mclip_pair_lse, mclip_merge_col_sums, mclip_lse_from_sum), so that the host logic around them -- the column-sum
    exchange, the status flag and the predicated fallback -- runs on CPU under gloo.  Arithmetic mirrors the kernels:
    f32 exponent range (values below 2^-126 flush to zero, above 2^128 overflow), sums in f64."""

    name = "oracle-emulation (two-sided forward)"

    def pair_supported(self, X, Y):
        return True

    def pair_ref(self, X, Y, ls, diag_off):
        self.calls.append("pair_ref")
        M, N = X.shape[0], Y.shape[0]
        idx = torch.arange(M) + diag_off
        ok = (idx >= 0) & (idx < N)
        diag = torch.zeros(M, dtype=torch.float64)
        diag[ok] = (X.double()[ok] * Y.double()[idx[ok]]).sum(dim=1)
        k2 = float(ls) * LOG2E
        c0 = float(torch.maximum(k2 * diag[ok].max(), k2 * diag[ok].min())) - 15.0 if bool(ok.any()) else 0.0
        return diag.float(), torch.tensor([c0], dtype=torch.float32), torch.zeros(1, dtype=torch.int32)

    def pair_lse(self, X, Y, ls, ref, status, want_rowdot, col_mode=0, diag=None, diag_off=0, out_msg=None, out_rowdot=None):
        self.calls.append("pair_lse")
        C = X.double() @ Y.double().T
        c0 = float(ref)
        t = (float(ls) * LOG2E * C - c0).clamp(max=127.99)
        e = torch.where(t < -126.0, torch.zeros_like(t), torch.exp2(t))          # f32 range of ex2.approx.ftz
        e = torch.where(t >= 127.99, torch.full_like(t, float("inf")), e)
        rs, cs = e.sum(dim=1), e.sum(dim=0)
        if not bool(((rs >= SUM_LO) & (rs <= SUM_HI)).all()):
            status |= 1
        row_lse = ((c0 + torch.log2(rs)) * LN2).float()
        rowdot = ((e * C).sum(dim=1) / rs).float() if want_rowdot else None
        if col_mode == 0:
            if not bool(((cs >= SUM_LO) & (cs <= SUM_HI)).all()):
                status |= 2
            return row_lse, rowdot, ((c0 + torch.log2(cs)) * LN2).float()
        if not bool((cs <= SUM_HI).all()):
            status |= 2
        N = Y.shape[0]
        out = torch.empty(N + 2, dtype=torch.float32) if out_msg is None else out_msg[:N + 2]
        out[:N] = cs.float()
        out[N] = c0
        out[N + 1:N + 2] = status.view(torch.float32)
        if out_msg is not None:                       # single-message form: row LSEs ride behind the column vector
            out_msg[N + 2:] = row_lse
            row_lse = out_msg[N + 2:]
            if want_rowdot and out_rowdot is not None:
                out_rowdot.copy_(rowdot)
                rowdot = out_rowdot
        return row_lse, rowdot, out

    def merge_col_sums(self, parts, n_total, col0, n, status):
        self.calls.append("merge_col_sums")
        refs = parts[:, n_total].double()
        bits = parts[:, n_total + 1].contiguous().view(torch.int32)
        for b in bits.tolist():
            status |= b
        m = refs.max()
        tot = (parts[:, col0:col0 + n].double() * torch.exp2(refs - m)[:, None]).sum(dim=0)
        if not bool(((tot >= SUM_LO) & (tot <= SUM_HI)).all()):
            status |= 2
        return ((m + torch.log2(tot)) * LN2).float()

    def lse_from_sum(self, sums, ref, status):
        self.calls.append("lse_from_sum")
        s = sums.double()
        if not bool(((s >= SUM_LO) & (s <= SUM_HI)).all()):
            status |= 2
        return ((float(ref) + torch.log2(s)) * LN2).float()
