"""Producer epilogue in front of the loss.

`ClipModel.encode_image/encode_text` (reference model.py:1011-1017) end in `F.normalize(features, dim=-1)`; under AMP
the fp32 result is then rounded to the autocast dtype by the logits matmul.  `normalize_features` does both in one
pass (and one pass backward), handing `ClipLoss` 16-bit unit-norm features that it can feed to the tensor cores
directly."""
from __future__ import annotations

import torch

from . import _cabi
from ._function import ClipLossFromProjectionsFunction, _check_equal_shards


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


def clip_loss_from_projections(loss_module, image_projections: torch.Tensor, text_projections: torch.Tensor, logit_scale,
                               dtype: torch.dtype = torch.bfloat16, eps: float = 1e-12, output_dict: bool = True):
    """The tail of `ClipModel.forward` + `ClipLoss.forward` in one call: `F.normalize(proj, dim=-1)` of both towers
    (reference model.py:1013,1017), the rounding to the autocast dtype in front of the logits matmul, the feature gather
    (loss.py:16-44) and the loss (loss.py:124-147).  The normalise+cast kernel writes each rank's shard directly into its
    slot of the all-gather buffer (the gather then runs in place), so the raw fp32 projections are read exactly once.

    `loss_module` is the training loop's `ClipLoss` (its mode flags, rank, world_size and group are used)."""
    if image_projections.dim() != 2 or image_projections.shape != text_projections.shape:
        raise ValueError(f"projections must be two [B, D] tensors of equal shape; got {tuple(image_projections.shape)} and "
                         f"{tuple(text_projections.shape)}")
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        raise ValueError(f"unsupported compute dtype {dtype}")
    W = int(loss_module.world_size)
    group = getattr(loss_module, "group", None)
    if W > 1:
        _check_equal_shards(image_projections, group)
    total = ClipLossFromProjectionsFunction.apply(image_projections, text_projections, logit_scale, dtype, float(eps),
                                                  bool(loss_module.local_loss), bool(loss_module.gather_with_grad),
                                                  int(loss_module.rank), W, group)
    return {"contrastive_loss": total} if output_dict else total
