"""Drop-in replacement for `mamba_clip.loss` (reference src/mamba_clip/loss.py).

Same public names, constructor order, attributes and call contract, so that
`from mamba_clip_b200.loss import ClipLoss, create_loss, all_gather, cross_entropy_loss` can stand in at
reference pipeline.py:15,546-552, integrations/ray.py:20,94-100 and is driven unchanged by
train.py:189 (`loss(**model_out, target=targets)`).  The arithmetic of `ClipLoss.forward` runs in
hand-written sm_100a kernels through the C ABI of include/mclip_b200.h; the B x B logits are never
materialised.  There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from ._function import _gather_rows, clip_loss


def create_loss(args):
    """reference loss.py:6-13."""
    return ClipLoss(
        local_loss=args.local_loss,
        gather_with_grad=args.gather_with_grad,
        cache_labels=True,
        rank=args.rank,
        world_size=args.world_size,
    )


def all_gather(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1):
    """Feature gather with the reference's semantics (loss.py:16-44), into contiguous buffers.

    `ClipLoss.forward` does not use this (its gather lives in the autograd Function); it is kept because
    it is part of the module's public surface.  With `gather_with_grad` the result is differentiable
    (backward = reduce-scatter SUM, as torch's `_AllGather`); otherwise only the rank's own slot carries
    grad, and only when `not local_loss`.
    """
    if gather_with_grad:
        import torch.distributed.nn.functional as dist_fn  # the reference forgets this import (loss.py:26)
        all_image = torch.cat(dist_fn.all_gather(image_features), dim=0)
        all_text = torch.cat(dist_fn.all_gather(text_features), dim=0)
        return all_image, all_text
    with torch.no_grad():
        all_image = _gather_rows(image_features, world_size, None)
        all_text = _gather_rows(text_features, world_size, None)
    if not local_loss:
        n = image_features.shape[0]
        pieces_i = list(all_image.split(n, dim=0))
        pieces_t = list(all_text.split(n, dim=0))
        pieces_i[rank] = image_features
        pieces_t[rank] = text_features
        all_image = torch.cat(pieces_i, dim=0)
        all_text = torch.cat(pieces_t, dim=0)
    return all_image, all_text


def cross_entropy_loss(input: torch.Tensor, target: torch.Tensor, weight=None) -> torch.Tensor:
    """Stage-2 classifier loss (reference loss.py:47-53); [B, 2] logits, not on the hot path."""
    if target.dtype in (torch.float, torch.double):
        return -(input.log_softmax(dim=-1) * target).sum(dim=-1).mean()
    return F.cross_entropy(input, target, weight=weight)


class ClipLoss(torch.nn.Module):
    """CLIP / InfoNCE loss over features gathered from every rank (reference loss.py:56-147).

    No parameters or buffers (checkpoints are unaffected).  `forward` returns
    `{"contrastive_loss": loss}` by default, like the reference (`output_dict=True`).

    Input contract (the producer's, reference model.py:1011-1017): features are L2-normalised rows.  Arbitrary finite
    features are accepted and the loss value is exact for them, but for **bf16** inputs the backward multiplies with an f16
    copy of the features (exact for 6.1e-5 <= |v| <= 65504; saturating above), so bf16 features beyond +-65504 would see
    clamped gradients.  fp16 and fp32 inputs have no such restriction.  `backward` is once-differentiable.
    """

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        # cache state (kept for API parity; the fused kernels take the label offset as an integer)
        self.prev_num_logits = 0
        self.labels = {}
        # optional process group (None = default group, as the reference)
        self.group = None

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        """reference loss.py:76-87: arange(n) (+ n*rank for local_loss on W>1), cached per device."""
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        """Materialised logits (reference loss.py:89-113).  Debug utility for small sizes only -- the
        training path (`forward`) never forms these matrices."""
        if self.world_size > 1:
            all_image, all_text = all_gather(image_features, text_features, self.local_loss,
                                             self.gather_with_grad, self.rank, self.world_size)
            if self.local_loss:
                logits_per_image = logit_scale * image_features @ all_text.T
                logits_per_text = logit_scale * text_features @ all_image.T
            else:
                logits_per_image = logit_scale * all_image @ all_text.T
                logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        return logits_per_image, logits_per_text

    def _gather_labels(self, labels):
        """reference loss.py:115-122 (dead code there; kept for API parity)."""
        if self.world_size > 1:
            gathered = [torch.zeros_like(labels) for _ in range(self.world_size)]
            dist.all_gather(gathered, labels)
            if not self.local_loss:
                gathered[self.rank] = labels
            return torch.cat(gathered, dim=0)
        return labels

    def forward(self, image_features, text_features, logit_scale, output_dict=True, target=None):
        # `target` is accepted and ignored, as in the reference (loss.py:137-140)
        total_loss = clip_loss(image_features, text_features, logit_scale, self.local_loss, self.gather_with_grad,
                               self.rank, self.world_size, self.group)
        if self.cache_labels:  # keep the observable cache state in step with the reference
            n = image_features.shape[0] * (1 if (self.world_size == 1 or self.local_loss) else self.world_size)
            self.prev_num_logits = n
        return {"contrastive_loss": total_loss} if output_dict else total_loss
