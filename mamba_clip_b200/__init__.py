"""B200-native (sm_100a) CLIP contrastive loss: drop-in for `mamba_clip.loss` of psmyth94/mamba-clip."""
from ._function import cuda_graphs_enabled, enable_cuda_graphs  # noqa: F401
from .loss import ClipLoss, all_gather, create_loss, cross_entropy_loss  # noqa: F401

__all__ = ["ClipLoss", "all_gather", "create_loss", "cross_entropy_loss", "enable_cuda_graphs", "cuda_graphs_enabled"]
__version__ = "0.1.0"
