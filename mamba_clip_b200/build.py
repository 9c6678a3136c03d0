"""Build libmclip_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mamba_clip_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmclip_b200.so")
SOURCES = ["abi.cu", "simt_kernels.cu", "small_kernels.cu", "tc_kernels.cu", "tc_bwd_persist.cu", "tc_pair_lse.cu", "tc_gemm_tn.cu", "producer_kernels.cu"]
HEADERS = ["common.cuh", "sm100_ptx.cuh", "tc_host.cuh", "tc_bwd_common.cuh", os.path.join("..", "..", "include", "mclip_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


PROFILE_LIB = os.path.join(HERE, "libmclip_b200_prof.so")


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    """Release library (default) or, with `profile`, a separate -DMCLIP_PROFILE library next to it (in-kernel cycle
    accounting, selected at run time with MCLIP_LIB_PATH=.../libmclip_b200_prof.so MCLIP_DBG=16)."""
    out = PROFILE_LIB if profile else LIB
    if not force and not profile and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    if profile:
        cmd.insert(1, "-DMCLIP_PROFILE")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--profile" in sys.argv, verbose="-v" in sys.argv,
                profile="--profile" in sys.argv))
