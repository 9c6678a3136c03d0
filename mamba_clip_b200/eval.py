"""Eval-time contrastive loss.

`evaluate` (reference eval.py:107-116) inlines a per-batch, single-process contrastive loss:
`logits = logit_scale.mean() * image_features @ text_features.t()`, labels `arange(batch)`, mean of the two cross
entropies.  This is the forward half of `ClipLoss` at world_size 1; here it runs on the fused kernels without
materialising the logits (the reference carries a FIXME about exactly that, eval.py:65-66)."""
from __future__ import annotations

import torch

from ._function import clip_loss


@torch.no_grad()
def contrastive_eval_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale) -> torch.Tensor:
    """Forward-only loss of one validation batch (no gather, no gradient); returns a 0-dim f32 tensor."""
    ls = logit_scale.mean() if torch.is_tensor(logit_scale) else float(logit_scale)
    return clip_loss(image_features, text_features, ls, False, False, 0, 1, None)
