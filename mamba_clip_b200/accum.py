"""Gradient-accumulation form of the contrastive loss.

The reference's `accum_freq > 1` path (train.py:198-290) first caches the features of `accum_freq` micro-batches without
gradient, then re-runs the model on micro-batch j and means to score it against the cached features of the others:
`inputs[key] = torch.cat(accumulated[:j] + [model_out[key]] + accumulated[j + 1:])` (train.py:262-270) -- as shipped the
call that follows passes `model_out` instead of `inputs` (train.py:272), so the path never ran; this module implements
the intended semantics.  With a loss that never materialises the logits the concatenated-negatives formulation is cheap:
the forward is one pass over the `[accum*B_l, accum*B_g]` block, and since only micro-batch j carries gradient the two
backward recompute launches cover just its rows (1/accum_freq of the backward work of calling `ClipLoss` on the
concatenation, same result)."""
from __future__ import annotations

from typing import Sequence

import torch

from ._function import ClipLossChunkFunction


def clip_loss_accum(loss_module, accum_image_features: Sequence[torch.Tensor], accum_text_features: Sequence[torch.Tensor],
                    j: int, image_features: torch.Tensor, text_features: torch.Tensor, logit_scale,
                    output_dict: bool = True):
    """Loss of micro-batch `j` (live, with gradient) against the cached features of the other micro-batches.

    `loss_module` is the `ClipLoss` instance of the training loop (its `local_loss`, `gather_with_grad`, `rank`,
    `world_size` are used).  Equivalent to
    `loss_module(torch.cat(acc_i[:j] + [image_features] + acc_i[j+1:]), torch.cat(acc_t[:j] + [text_features] + acc_t[j+1:]), logit_scale)`.
    """
    n = len(accum_image_features)
    if n != len(accum_text_features) or not (0 <= j < n):
        raise ValueError(f"need equally long feature caches and 0 <= j < {n}; got {len(accum_text_features)} text caches, j={j}")
    if image_features.shape != accum_image_features[j].shape or text_features.shape != accum_text_features[j].shape:
        raise ValueError("the live micro-batch must have the shape of the cached micro-batch it replaces")
    if image_features.shape != text_features.shape or image_features.dim() != 2:
        raise ValueError(f"features must be [B, D] and equal in shape; got {tuple(image_features.shape)}, {tuple(text_features.shape)}")
    with torch.no_grad():
        full_i = torch.cat([f.detach() for f in accum_image_features], dim=0)
        full_t = torch.cat([f.detach() for f in accum_text_features], dim=0)
    lo = sum(int(f.shape[0]) for f in accum_image_features[:j])
    total = ClipLossChunkFunction.apply(image_features, text_features, logit_scale, full_i, full_t, lo,
                                        bool(loss_module.local_loss), bool(loss_module.gather_with_grad),
                                        int(loss_module.rank), int(loss_module.world_size), getattr(loss_module, "group", None))
    return {"contrastive_loss": total} if output_dict else total
