"""Loop-level drop-in check: the reference's UNCHANGED `train_one_epoch` (train.py:92-385) driven with our
ClipLoss vs the reference ClipLoss, same seeds, same stand-in model -> same parameter trajectory.

CPU only (the C-ABI primitives are replaced by their oracle statements) and only where /root/reference exists
(the build container); the GPU box has no copy of the reference."""
import copy
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


class Towers(torch.nn.Module):
    """Dict-output stand-in with the contract of ClipModel.forward (reference model.py:1047-1057)."""

    def __init__(self, dim=32):
        super().__init__()
        self.visual = torch.nn.Linear(3 * 8 * 8, dim)
        self.text = torch.nn.Embedding(50, dim)
        self.logit_scale = torch.nn.Parameter(torch.tensor(np.log(1 / 0.07), dtype=torch.float32))

    def forward(self, image, text):
        i = torch.nn.functional.normalize(self.visual(image.flatten(1)), dim=-1)
        t = torch.nn.functional.normalize(self.text(text).mean(1), dim=-1)
        return {"image_features": i, "text_features": t, "logit_scale": self.logit_scale.exp()}


class Loader:
    def __init__(self, batches):
        self.batches = batches
        self.num_batches = len(batches)
        self.num_samples = len(batches) * batches[0][0][0].shape[0]

    def __iter__(self):
        return iter(self.batches)


class Data:
    def __init__(self, loader):
        self.dataloader = loader

    def set_epoch(self, epoch):
        pass


def make_batches(n, bs):
    g = torch.Generator().manual_seed(5)
    out = []
    for _ in range(n):
        def one():
            return (torch.randn(bs, 3, 8, 8, generator=g), torch.randint(0, 50, (bs, 6), generator=g),
                    torch.randint(0, 2, (bs,), generator=g))
        out.append([one(), one()])     # ComboLoader shape: [batch, balanced_batch] (reference data.py:218-239)
    return out


def run_epoch(loss_module):
    sys.path.insert(0, REF)
    import torch.distributed.nn  # noqa: F401  (reference loss.py:26 relies on this import)
    from mamba_clip.train import train_one_epoch
    torch.manual_seed(0)
    np.random.seed(0)
    model = Towers()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    args = types.SimpleNamespace(device="cpu", precision="fp32", accum_freq=1, skip_scheduler=True, balanced_mixup=0.4,
                                 num_classes=2, grad_clip_norm=None, log_every_n_steps=1, batch_size=16, world_size=1,
                                 wandb=False, rank=0, local_rank=0, hyperparameter_tuning=False)
    data = {"train": Data(Loader(make_batches(3, 16)))}
    train_one_epoch(model, data, loss_module, 0, opt, None, None, args, tb_writer=None)
    return copy.deepcopy(model.state_dict())


def test_reference_train_loop_accepts_the_dropin():
    sys.path.insert(0, REF)
    from mamba_clip.loss import ClipLoss as RefLoss
    from mamba_clip_b200 import ClipLoss, _cabi
    from tests._emul import EmulatedBackend
    ref_state = run_epoch(RefLoss(cache_labels=True))
    _cabi.set_backend_override(EmulatedBackend())
    try:
        new_state = run_epoch(ClipLoss(cache_labels=True))
    finally:
        _cabi.set_backend_override(None)
    for k in ref_state:
        a, b = ref_state[k].double(), new_state[k].double()
        assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max())), k
