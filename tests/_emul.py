"""Oracle-backed stand-in for the four C-ABI primitives, for CPU tests of the host logic only
(world_size-2/4 gloo).  Lives under tests/: the product never imports it."""
import torch

from oracle import clip_oracle as O


class EmulatedBackend:
    name = "oracle-emulation"

    def __init__(self):
        self.calls = []

    def row_lse(self, X, Y, ls, diag_off, want_diag, want_rowdot=False):
        self.calls.append("row_lse")
        lse, diag = O.block_row_lse(X, Y, float(ls), diag_off)
        out = (lse.float(), (diag.float() if want_diag else None))
        if want_rowdot:
            C = X.double() @ Y.double().T
            P = torch.exp(float(ls) * C - lse[:, None])
            out = out + ((P * C).sum(dim=1).float(),)
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
        t = (u.double() + v.double() - 2 * diag.double()).sum()
        return t.float(), (float(go) * scale * t).float()

    def launch_count(self):
        return len(self.calls)
