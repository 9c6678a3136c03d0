import os
import sys

import pytest

# tests (and only tests) may install an oracle-backed stand-in for the C-ABI primitives to run the multi-rank host logic
# under gloo on CPU: mamba_clip_b200._cabi.set_backend_override refuses without this opt-in
os.environ.setdefault("MCLIP_ALLOW_TEST_BACKEND", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no CUDA device is visible, so that a bare `pytest tests/`
    # on the CPU box still passes; `-m gpu` on the GPU box runs them for real.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
