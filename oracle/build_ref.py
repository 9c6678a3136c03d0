"""Recipe for oracle/_ref/: the UNMODIFIED reference implementation of the hot path.  TEST / BASELINE INFRASTRUCTURE ONLY.

    python -m oracle.build_ref            # in the build container, where /root/reference exists

The reference's ClipLoss lives in one self-contained Python file (src/mamba_clip/loss.py: torch, torch.distributed and
torch.nn.functional are its only imports), so "building" it is a byte-for-byte copy to
`oracle/_ref/mamba_clip_loss_ref.py` plus a sha256 manifest.  `oracle/_ref/` is git-ignored (no reference source enters
the history) but travels to the GPU box with the gpurun snapshot, where `bench.py --impl reference` / the `cpu_baseline`
leg time it on the host cores and `tests/test_oracle.py` re-validates the port against it.  Nothing in the product
(`mamba_clip_b200/`) may import it.

`load()` returns the module with the one shim the survey documents: `import torch.distributed.nn` before use
(the reference calls `dist.nn.all_gather` without importing the sub-module, loss.py:26).
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_FILE = os.path.join(REF_DIR, "mamba_clip_loss_ref.py")
MANIFEST = os.path.join(REF_DIR, "MANIFEST.json")
SOURCE = os.environ.get("MCLIP_REFERENCE_LOSS", "/root/reference/src/mamba_clip/loss.py")


def _sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(force: bool = False) -> str | None:
    """Copy the reference file unchanged.  Returns the path, or None when the reference tree is absent (GPU box)."""
    if not os.path.exists(SOURCE):
        return REF_FILE if os.path.exists(REF_FILE) else None
    if os.path.exists(REF_FILE) and not force and _sha256(REF_FILE) == _sha256(SOURCE):
        return REF_FILE
    os.makedirs(REF_DIR, exist_ok=True)
    shutil.copyfile(SOURCE, REF_FILE)
    with open(MANIFEST, "w") as f:
        json.dump({"source": SOURCE, "sha256": _sha256(REF_FILE), "bytes": os.path.getsize(REF_FILE),
                   "note": "byte-for-byte copy of the reference loss.py; not tracked by git"}, f, indent=1)
    return REF_FILE


def available() -> bool:
    return os.path.exists(REF_FILE)


def load():
    """Import oracle/_ref/mamba_clip_loss_ref.py (the unmodified reference loss.py) as a module."""
    if not available():
        raise FileNotFoundError(f"{REF_FILE} missing: run `python -m oracle.build_ref` where /root/reference exists")
    import torch.distributed.nn  # noqa: F401  -- the shim: reference loss.py:26 uses dist.nn without importing it
    name = "mamba_clip_loss_ref"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree not present; nothing built")
