"""Producer epilogue in front of the loss.

`ClipModel.encode_image/encode_text` (reference model.py:1011-1017) end in `F.normalize(features, dim=-1)`; under AMP
the fp32 result is then rounded to the autocast dtype by the logits matmul.  `normalize_features` does both in one
pass (and one pass backward), handing `ClipLoss` 16-bit unit-norm features that it can feed to the tensor cores
directly."""
from __future__ import annotations

import torch

from . import _cabi


class _NormalizeCast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_dtype, eps):
        be = _cabi.get_backend()
        xf = x.detach().to(torch.float32).contiguous()
        y = be.normalize_rows(xf, out_dtype, eps)
        ctx.save_for_backward(xf)
        ctx.eps = eps
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        (xf,) = ctx.saved_tensors
        dx = _cabi.get_backend().normalize_rows_bwd(xf, g.contiguous(), ctx.eps)
        return dx.to(ctx.in_dtype), None, None


def normalize_features(features: torch.Tensor, dtype: torch.dtype = torch.bfloat16, eps: float = 1e-12) -> torch.Tensor:
    """`F.normalize(features, dim=-1)` (reference model.py:1013,1017) fused with the cast to `dtype`; differentiable."""
    if features.dim() != 2:
        raise ValueError(f"features must be [B, D]; got {tuple(features.shape)}")
    if dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise ValueError(f"unsupported output dtype {dtype}")
    return _NormalizeCast.apply(features, dtype, float(eps))
