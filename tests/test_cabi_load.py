"""The C-ABI shared library: builds for sm_100a, loads without a GPU, exports every symbol that
include/mclip_b200.h declares, and rejects bad arguments with error codes (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mamba_clip_b200 import _cabi, build
    build.build()
    return _cabi.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mclip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mclip_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    from mamba_clip_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mclip_b200.h but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == names
    assert lib.mclip_abi_version() == _cabi.ABI_VERSION


def test_no_torch_or_libcuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "mamba_clip_b200", "libmclip_b200.so")],
                         capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    assert "libcuda.so" not in out  # driver entry points are resolved at run time


def test_argument_validation_without_gpu(lib):
    n = ctypes.c_size_t(0)
    assert lib.mclip_workspace_bytes(128, 128, 64, 1, 0, 0, ctypes.byref(n)) == 0
    assert lib.mclip_workspace_bytes(0, 128, 64, 1, 0, 0, ctypes.byref(n)) == 1
    assert lib.mclip_workspace_bytes(128, 128, 64, 7, 0, 0, ctypes.byref(n)) == 1
    assert b"invalid" in lib.mclip_last_error()
    # null pointers -> MCLIP_ERR_INVALID before any CUDA call
    rc = lib.mclip_row_lse(None, None, 4, 4, 8, 8, 8, 0, None, 0, None, None, None, None, None, 0, 0, None)
    assert rc == 1 and b"null" in lib.mclip_last_error()
    rc = lib.mclip_block_grad(None, None, 4, 4, 8, 8, 8, 0, None, None, None, None, 0, 1.0, 1.0, 2.0, 0.5,
                              None, 8, None, None, None, 0, 0, None)
    assert rc == 1
    assert lib.mclip_convert_f16(None, 4, 8, 8, None, None) == 1
    assert lib.mclip_loss_finalize(None, None, None, 4, None, None, None) == 1
    assert lib.mclip_dls_finalize(None, None, None, 4, None, 1.0, None, None, None) == 1
    # two-sided forward entry points
    assert lib.mclip_pair_ref(None, None, 4, 4, 8, 8, 8, 1, None, 0, None, None, None, None, 0, None) == 1
    assert lib.mclip_pair_lse(None, None, 4, 4, 8, 8, 8, 1, None, None, 0, None, None, None, None, 0, None, None, 0, None) == 1
    assert lib.mclip_merge_col_sums(None, 2, 10, 8, 0, 4, None, None, None) == 1
    assert lib.mclip_lse_from_sum(None, 4, None, None, None, None) == 1
    assert lib.mclip_pair_supported(4096, 4096, 512, 512, 512, 1) == 1        # bf16, D <= 512
    assert lib.mclip_pair_supported(4096, 4096, 768, 768, 768, 1) == 1        # 512 < D <= 768: X k-chunks past 512 streamed
    assert lib.mclip_pair_supported(4096, 4096, 832, 832, 832, 1) == 0        # D > 768: one-sided kernels
    # round-2 entry points: argument validation, support queries, options (no GPU work)
    assert lib.mclip_fused_grad(None, None, 4096, 8192, 512, 512, 512, 1, None, None, None, None, 0, 0.5, None, 512, None, 512,
                                None, None, 0, None) == 1
    assert lib.mclip_small_supported(64, 512, 512, 1) == 1 and lib.mclip_small_supported(64, 2048, 512, 1) == 0
    assert lib.mclip_small_supported(64, 512, 520, 0) == 0 and lib.mclip_small_supported(64, 64, 512, 0) == 1
    assert lib.mclip_small_counter_words(64, 512) >= 17
    assert lib.mclip_small_forward(None, None, 64, 512, 512, 65536, 1, None, 0, 512, None, None, 0, None, None) == 1
    assert lib.mclip_small_pack(None, None, 10, 1, 1, None, None) == 1
    assert lib.mclip_set_option(b"no_such_option", 1) == 1
    import ctypes as _ct
    v, old = _ct.c_int(-7), _ct.c_int(-7)
    assert lib.mclip_get_option(b"bwd_persist", _ct.byref(old)) == 0 and old.value in (-1, 0, 1)
    assert lib.mclip_set_option(b"bwd_persist", 1) == 0 and lib.mclip_get_option(b"bwd_persist", _ct.byref(v)) == 0 and v.value == 1
    assert lib.mclip_set_option(b"bwd_persist", old.value) == 0
    assert lib.mclip_pair_supported(4096, 4096, 512, 512, 512, 0) == 0        # fp32: FFMA path
    for op in (2, 3):                                                          # PAIR_LSE, PAIR_REF workspaces
        assert lib.mclip_workspace_bytes(32768, 32768, 512, 1, op, 0, ctypes.byref(n)) == 0 and n.value > 0
    # the measurement hook can be read without a GPU (nothing recorded)
    tot, cnt = ctypes.c_float(-1.0), ctypes.c_int(-1)
    assert lib.mclip_kernel_timing(0, ctypes.byref(tot), ctypes.byref(cnt)) == 0 and cnt.value == 0 and tot.value == 0.0


def test_path_selection(lib):
    TC, SIMT = 2, 1
    assert lib.mclip_select_path(4096, 4096, 512, 512, 512, 1, 0) == TC      # bf16, D=512
    assert lib.mclip_select_path(4096, 4096, 512, 512, 512, 2, 1) == TC      # f16
    assert lib.mclip_select_path(4096, 4096, 512, 512, 512, 0, 0) == SIMT    # fp32 needs fp32 math (1e-5 bar)
    assert lib.mclip_select_path(64, 64, 100, 100, 100, 1, 0) == SIMT        # D % 8 != 0
    assert lib.mclip_select_path(64, 64, 1024, 1024, 1024, 1, 0) == SIMT     # D > 768
    assert lib.mclip_select_path(64, 64, 768, 768, 768, 1, 0) == TC


def test_sass_has_blackwell_instructions():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mamba_clip_b200", "libmclip_b200.so")],
                          capture_output=True, text=True).stdout
    # tcgen05.mma (CTA pairs), TMA loads / stores / reduce-adds, tcgen05.ld; G is handed over through shared memory,
    # so there is no tcgen05.st (STTM) any more
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16816" not in sass  # no legacy mma.sync path
