"""ctypes binding of libmclip_b200.so (include/mclip_b200.h) + the tensor-level backend the loss uses.

There is no CPU implementation here on purpose: if the library is missing, or a tensor is not on a
CUDA device, the calls raise.  (tests/ may install an oracle-backed stand-in through
`set_backend_override` to exercise the host logic under gloo on CPU.)
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCLIP_LIB_PATH: development only (e.g. a -DMCLIP_PROFILE build next to the release library)
LIB_PATH = os.environ.get("MCLIP_LIB_PATH") or os.path.join(_HERE, "libmclip_b200.so")

ABI_VERSION = 6
DTYPE_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2
OP_ROW_LSE, OP_BLOCK_GRAD, OP_PAIR_LSE, OP_PAIR_REF, OP_FUSED_GRAD, OP_SMALL = 0, 1, 2, 3, 4, 5

_c_f32p = ctypes.c_void_p
_SIGNATURES = {
    "mclip_abi_version": (ctypes.c_int, []),
    "mclip_last_error": (ctypes.c_char_p, []),
    "mclip_launch_count": (ctypes.c_int64, []),
    "mclip_device_supported": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "mclip_select_path": (ctypes.c_int, [ctypes.c_int64] * 5 + [ctypes.c_int, ctypes.c_int]),
    "mclip_workspace_bytes": (ctypes.c_int, [ctypes.c_int64] * 3 + [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_size_t)]),
    "mclip_row_lse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5 + [ctypes.c_int, _c_f32p,
                                     ctypes.c_int64, _c_f32p, _c_f32p, _c_f32p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "mclip_pair_supported": (ctypes.c_int, [ctypes.c_int64] * 5 + [ctypes.c_int]),
    "mclip_pair_ref": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5 + [ctypes.c_int, _c_f32p,
                                      ctypes.c_int64, _c_f32p, _c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_void_p]),
    "mclip_pair_lse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5 + [ctypes.c_int, _c_f32p,
                                      _c_f32p, ctypes.c_int64, _c_f32p, _c_f32p, _c_f32p, _c_f32p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p]),
    "mclip_merge_col_sums": (ctypes.c_int, [_c_f32p, ctypes.c_int] + [ctypes.c_int64] * 4 + [_c_f32p, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "mclip_lse_from_sum": (ctypes.c_int, [_c_f32p, ctypes.c_int64, _c_f32p, _c_f32p, ctypes.c_void_p, ctypes.c_void_p]),
    "mclip_block_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5 + [ctypes.c_int, _c_f32p,
                                        _c_f32p, _c_f32p, _c_f32p, ctypes.c_int64] + [ctypes.c_float] * 4 +
                         [ctypes.c_void_p, ctypes.c_int64, _c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                          ctypes.c_void_p]),
    "mclip_convert_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "mclip_fused_grad_supported": (ctypes.c_int, [ctypes.c_int64] * 5 + [ctypes.c_int]),
    "mclip_fused_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5 + [ctypes.c_int, _c_f32p,
                                        _c_f32p, _c_f32p, _c_f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, _c_f32p, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p]),
    "mclip_small_supported": (ctypes.c_int, [ctypes.c_int64] * 3 + [ctypes.c_int]),
    "mclip_small_counter_words": (ctypes.c_int, [ctypes.c_int64] * 2),
    "mclip_small_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 4 + [ctypes.c_int, _c_f32p,
                                           ctypes.c_int64, ctypes.c_int64, _c_f32p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "mclip_small_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 4 + [ctypes.c_int, _c_f32p,
                                            _c_f32p, _c_f32p, ctypes.c_int64] + [ctypes.c_float] * 5 +
                             [ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                              ctypes.c_void_p]),
    "mclip_small_pack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "mclip_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "mclip_get_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]),
    "mclip_loss_finalize": (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, ctypes.c_int64, _c_f32p, _c_f32p, ctypes.c_void_p]),
    "mclip_dls_finalize": (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, ctypes.c_int64, _c_f32p, ctypes.c_float, _c_f32p,
                                          _c_f32p, ctypes.c_void_p]),
    "mclip_kernel_timing": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "mclip_normalize_rows": (ctypes.c_int, [_c_f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int64, _c_f32p, ctypes.c_void_p]),
    "mclip_normalize_rows_bwd": (ctypes.c_int, [_c_f32p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_int64, ctypes.c_int, ctypes.c_float, _c_f32p, ctypes.c_int64,
                                                ctypes.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the C-ABI library and bind every symbol the header declares.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -m mamba_clip_b200.build` (nvcc, sm_100a). "
            "mamba_clip_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.mclip_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libmclip_b200 ABI {got} != expected {ABI_VERSION}; rebuild the library")
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.mclip_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == 1 else RuntimeError
        raise exc(f"{what} failed (code {rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class CudaBackend:
    """Tensor-level wrapper: allocates outputs with torch, passes raw pointers + the current CUDA stream to the C
    ABI.  No host synchronisation anywhere.  Host overhead matters for the small, latency-bound configurations
    (B_g = 512), so workspace sizes are cached and one scratch buffer per (device, stream) is reused."""

    name = "cuda"

    def __init__(self, path: int = PATH_AUTO):
        self.lib = load_library()
        self.path = path
        self._checked = set()
        self._ws_size = {}
        self._ws_buf = {}
        self._counters = {}
        self._small_ok = {}

    # -- helpers ---------------------------------------------------------------------------------
    def _prep(self, *tensors):
        dev = tensors[0].device
        if dev.type != "cuda":
            raise RuntimeError("mamba_clip_b200 kernels need CUDA tensors (no CPU fallback); got " + str(dev))
        if dev.index not in self._checked:
            sm = ctypes.c_int(0)
            _check(self.lib, self.lib.mclip_device_supported(dev.index if dev.index is not None else torch.cuda.current_device(),
                                                             ctypes.byref(sm)), "mclip_device_supported")
            self._checked.add(dev.index)
        for t in tensors:
            if t is not None and t.device != dev:
                raise ValueError("all tensors must be on the same device")
        return dev

    def _workspace(self, M, N, D, dtype, op, dev, stream_ptr):
        key = (M, N, D, dtype, op)
        need = self._ws_size.get(key)
        if need is None:
            n = ctypes.c_size_t(0)
            _check(self.lib, self.lib.mclip_workspace_bytes(M, N, D, DTYPE_CODES[dtype], op, self.path, ctypes.byref(n)),
                   "mclip_workspace_bytes")
            need = self._ws_size[key] = n.value
        if need == 0:
            return None, 0
        # the library only uses the scratch inside the launches of one call, all on `stream_ptr`: reuse is stream-ordered
        bkey = (dev.index, stream_ptr)
        buf = self._ws_buf.get(bkey)
        if buf is None or buf.numel() < need:
            buf = self._ws_buf[bkey] = torch.empty(need, dtype=torch.uint8, device=dev)
        return buf, need

    def kernel_timing(self, enable: bool):
        """enable=True: start bracketing every launch of the dominant (pair backward) kernel with CUDA events.
        enable=False: stop and return (total_ms, launches) of what was recorded."""
        if enable:
            _check(self.lib, self.lib.mclip_kernel_timing(1, None, None), "mclip_kernel_timing")
            return None
        tot, n = ctypes.c_float(0.0), ctypes.c_int(0)
        _check(self.lib, self.lib.mclip_kernel_timing(0, ctypes.byref(tot), ctypes.byref(n)), "mclip_kernel_timing")
        return tot.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.mclip_launch_count())

    class _DeviceGuard:
        """`with torch.cuda.device(dev)` only when `dev` is not already current (the context manager is slow)."""
        def __init__(self, dev):
            self.ctx = None if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

        def __enter__(self):
            if self.ctx is not None:
                self.ctx.__enter__()

        def __exit__(self, *a):
            if self.ctx is not None:
                self.ctx.__exit__(*a)

    # -- ops -------------------------------------------------------------------------------------
    def row_lse(self, X: torch.Tensor, Y: torch.Tensor, ls: torch.Tensor, diag_off: int, want_diag: bool,
                want_rowdot: bool = False, run_if: Optional[torch.Tensor] = None, out_lse: Optional[torch.Tensor] = None,
                out_rowdot: Optional[torch.Tensor] = None):
        """-> (lse, diag or None) or, with want_rowdot, (lse, diag or None, rowdot).
        `run_if` (device int32 scalar): the launches exit without writing when it is 0 at execution time; such a
        call writes into the caller's `out_lse` / `out_rowdot` and cannot produce `diag`."""
        dev = self._prep(X, Y, ls)
        M, D = X.shape
        N = Y.shape[0]
        if run_if is not None:
            if want_diag or out_lse is None or (want_rowdot and out_rowdot is None):
                raise ValueError("a predicated row_lse needs caller-provided outputs and cannot write diag")
            lse, diag, rowdot = out_lse, None, (out_rowdot if want_rowdot else None)
        else:
            nout = 1 + int(want_diag) + int(want_rowdot)
            out = torch.empty((nout, M), dtype=torch.float32, device=dev)
            lse = out[0]
            diag = out[1] if want_diag else None
            rowdot = out[nout - 1] if want_rowdot else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(M, N, D, X.dtype, OP_ROW_LSE, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_row_lse(_ptr(X), _ptr(Y), M, N, D, X.stride(0), Y.stride(0), DTYPE_CODES[X.dtype],
                                        _ptr(ls), diag_off, _ptr(lse), _ptr(diag), _ptr(rowdot), _ptr(run_if), _ptr(ws),
                                        nws, self.path, ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_row_lse")
        return (lse, diag, rowdot) if want_rowdot else (lse, diag)

    def pair_supported(self, X: torch.Tensor, Y: torch.Tensor) -> bool:
        """True when the two-sided forward (one pass for both loss directions) can run this problem."""
        if self.path == PATH_SIMT or X.dtype not in DTYPE_CODES or os.environ.get("MCLIP_NO_PAIR_FWD") == "1":
            return False
        return bool(self.lib.mclip_pair_supported(X.shape[0], Y.shape[0], X.shape[1], X.stride(0), Y.stride(0),
                                                  DTYPE_CODES[X.dtype]))

    def pair_ref(self, X, Y, ls, diag_off):
        """-> (diag [M], ref [1], status [1] int32 zeroed): positive-pair dots and the uniform exponent reference."""
        dev = self._prep(X, Y, ls)
        M, D = X.shape
        N = Y.shape[0]
        out = torch.empty(M + 2, dtype=torch.float32, device=dev)
        diag, ref, status = out[:M], out[M:M + 1], out[M + 1:M + 2].view(torch.int32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(M, N, D, X.dtype, OP_PAIR_REF, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_pair_ref(_ptr(X), _ptr(Y), M, N, D, X.stride(0), Y.stride(0), DTYPE_CODES[X.dtype], _ptr(ls),
                                         diag_off, _ptr(diag), _ptr(ref), _ptr(status), _ptr(ws), nws, ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_pair_ref")
        return diag, ref, status

    def pair_lse(self, X, Y, ls, ref, status, want_rowdot: bool, col_mode: int = 0, diag=None, diag_off: int = 0,
                 out_msg: Optional[torch.Tensor] = None, out_rowdot: Optional[torch.Tensor] = None):
        """-> (row_lse [M], rowdot [M] or None, col_out [N]); ORs into `status` when the result is unusable.
        `diag` (from pair_ref, same diag_off) is overwritten in place with the accumulator's own positive-pair dots.
        `out_msg` (col_mode 1 only): one caller-provided f32 buffer [N + 2 + M] that receives the column vector
        (sums, ref, status) followed by the row LSEs -- the single message a rank all-gathers; `out_rowdot`: where to
        put rowdot [M]."""
        dev = self._prep(X, Y, ls, ref, status)
        M, D = X.shape
        N = Y.shape[0]
        if out_msg is not None:
            if col_mode != 1 or out_msg.numel() != N + 2 + M or out_msg.dtype != torch.float32 or not out_msg.is_contiguous():
                raise ValueError("out_msg needs col_mode=1 and a contiguous f32 buffer of N + 2 + M elements")
            col_out, row_lse = out_msg[:N + 2], out_msg[N + 2:]
            rowdot = (out_rowdot if out_rowdot is not None else torch.empty(M, dtype=torch.float32, device=dev)) if want_rowdot else None
        else:
            rows = torch.empty((2, M), dtype=torch.float32, device=dev)
            col_out = torch.empty(N + (2 if col_mode == 1 else 0), dtype=torch.float32, device=dev)
            row_lse, rowdot = rows[0], (rows[1] if want_rowdot else None)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(M, N, D, X.dtype, OP_PAIR_LSE, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_pair_lse(_ptr(X), _ptr(Y), M, N, D, X.stride(0), Y.stride(0), DTYPE_CODES[X.dtype], _ptr(ls),
                                         _ptr(ref), diag_off, _ptr(diag), _ptr(row_lse), _ptr(rowdot), _ptr(col_out),
                                         col_mode, _ptr(status),
                                         _ptr(ws), nws, ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_pair_lse")
        return row_lse, rowdot, col_out

    def merge_col_sums(self, parts, n_total, col0, n, status):
        """parts: [W, n_total + 2] gathered col_mode-1 vectors -> column LSEs of columns [col0, col0 + n)."""
        dev = self._prep(parts, status)
        lse = torch.empty(n, dtype=torch.float32, device=dev)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_merge_col_sums(_ptr(parts), parts.shape[0], parts.stride(0), n_total, col0, n, _ptr(lse),
                                               _ptr(status), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_merge_col_sums")
        return lse

    def lse_from_sum(self, sums, ref, status):
        dev = self._prep(sums, ref, status)
        lse = torch.empty_like(sums)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_lse_from_sum(_ptr(sums), sums.numel(), _ptr(ref), _ptr(lse), _ptr(status),
                                             ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_lse_from_sum")
        return lse

    def to_f16(self, src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """f16 copy of a bf16 matrix (mclip_convert_f16) on the current stream; `out`: a [rows, D] contiguous f16 buffer."""
        dev = self._prep(src)
        rows, D = src.shape
        if out is None:
            out = torch.empty((rows, D), dtype=torch.float16, device=dev)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_convert_f16(_ptr(src), rows, D, src.stride(0), _ptr(out),
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_convert_f16")
        return out

    def block_grad(self, X, Y, ls, go, lse_x, lse_y, diag_off, w_row, w_col, w_diag, inv_2n, want_rowdot=True, y16=None):
        dev = self._prep(X, Y, ls, lse_x)
        M, D = X.shape
        N = Y.shape[0]
        dX = torch.empty((M, D), dtype=X.dtype, device=dev)
        rowdot = torch.empty(M, dtype=torch.float32, device=dev) if want_rowdot else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(M, N, D, X.dtype, OP_BLOCK_GRAD, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_block_grad(_ptr(X), _ptr(Y), M, N, D, X.stride(0), Y.stride(0), DTYPE_CODES[X.dtype],
                                           _ptr(ls), _ptr(go), _ptr(lse_x), _ptr(lse_y), diag_off, w_row, w_col,
                                           w_diag, inv_2n, _ptr(dX), dX.stride(0), _ptr(rowdot), _ptr(y16), _ptr(ws), nws,
                                           self.path, ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_block_grad")
        return dX, rowdot

    def fused_supported(self, X: torch.Tensor, Y: torch.Tensor) -> bool:
        """True when the shared-recompute backward (both gradients from one recompute of S) can run this problem and
        the problem is large enough to profit."""
        if self.path == PATH_SIMT or X.dtype not in DTYPE_CODES or X.dtype == torch.float32:
            return False
        return bool(self.lib.mclip_fused_grad_supported(X.shape[0], Y.shape[0], X.shape[1], X.stride(0), Y.stride(0),
                                                        DTYPE_CODES[X.dtype]))

    def fused_grad(self, X, Y, ls, go, lse_x, lse_y, diag_off, inv_2n):
        """-> (dX [M, D], dY [N, D], xdot [M]): mclip_fused_grad (weights (1, 1, 2), i.e. the full gradient)."""
        dev = self._prep(X, Y, ls, lse_x, lse_y)
        M, D = X.shape
        N = Y.shape[0]
        dX = torch.empty((M, D), dtype=X.dtype, device=dev)
        dY = torch.empty((N, D), dtype=X.dtype, device=dev)
        xdot = torch.empty(M, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(M, N, D, X.dtype, OP_FUSED_GRAD, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_fused_grad(_ptr(X), _ptr(Y), M, N, D, X.stride(0), Y.stride(0), DTYPE_CODES[X.dtype], _ptr(ls),
                                           _ptr(go), _ptr(lse_x), _ptr(lse_y), diag_off, inv_2n, _ptr(dX), dX.stride(0),
                                           _ptr(dY), dY.stride(0), _ptr(xdot), _ptr(ws), nws, ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_fused_grad")
        return dX, dY, xdot

    # -- latency path (small global batches): one forward and one backward kernel per step -----------------------
    def small_supported(self, Bl: int, Bg: int, D: int, dtype) -> bool:
        if self.path != PATH_AUTO or dtype not in DTYPE_CODES or os.environ.get("MCLIP_NO_SMALL_PATH") == "1":
            return False
        key = (Bl, Bg, D, dtype)
        ok = self._small_ok.get(key)
        if ok is None:       # this sits on the per-step host path of the latency configuration: ask the library once per shape
            ok = self._small_ok[key] = bool(self.lib.mclip_small_supported(Bl, Bg, D, DTYPE_CODES[dtype]))
        return ok

    def _small_counters(self, dev, stream_ptr):
        """Persistent zeroed counter words per (device, stream): the kernels leave them zeroed after every launch."""
        key = (dev.index, stream_ptr)
        c = self._counters.get(key)
        if c is None:
            c = self._counters[key] = torch.zeros(128, dtype=torch.int32, device=dev)
        return c

    def small_pack(self, a, b, out_dtype):
        """-> [2, Bl, D] contiguous send buffer (image shard, text shard) in the compute dtype."""
        dev = self._prep(a, b)
        out = torch.empty((2,) + tuple(a.shape), dtype=out_dtype, device=dev)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_small_pack(_ptr(a), _ptr(b), a.numel(), DTYPE_CODES[a.dtype], DTYPE_CODES[out_dtype], _ptr(out),
                                           ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_small_pack")
        return out

    def small_forward(self, A, B, Bl, Bg, D, blk_stride, ls, lo, hi):
        """-> stats [5 * Bg + 2] (row_lse, col_lse, diag, u, v, loss, t).  A / B: base tensors of the blocked layout."""
        dev = self._prep(A, B, ls)
        stats = torch.empty(5 * Bg + 2, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(Bl, Bg, D, A.dtype, OP_SMALL, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_small_forward(_ptr(A), _ptr(B), Bl, Bg, D, blk_stride, DTYPE_CODES[A.dtype], _ptr(ls), lo, hi,
                                              _ptr(stats), _ptr(ws), nws, _ptr(self._small_counters(dev, stream)),
                                              ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_small_forward")
        return stats

    def small_backward(self, A, B, Bl, Bg, D, blk_stride, ls, go, stats, off, w_row, w_col, w_diag, inv_2n, dls_scale):
        """-> (dA [Bl, D], dB [Bl, D], dls [1])."""
        dev = self._prep(A, B, ls, stats)
        out = torch.empty((2, Bl, D), dtype=A.dtype, device=dev)
        dls = torch.empty(1, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, nws = self._workspace(Bl, Bg, D, A.dtype, OP_SMALL, dev, stream)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_small_backward(_ptr(A), _ptr(B), Bl, Bg, D, blk_stride, DTYPE_CODES[A.dtype], _ptr(ls), _ptr(go),
                                               _ptr(stats), off, w_row, w_col, w_diag, inv_2n, dls_scale, _ptr(out[0]),
                                               _ptr(out[1]), _ptr(dls), _ptr(ws), nws,
                                               _ptr(self._small_counters(dev, stream)), ctypes.c_void_p(stream))
        _check(self.lib, rc, "mclip_small_backward")
        return out[0], out[1], dls

    def set_option(self, name: str, value: int) -> None:
        _check(self.lib, self.lib.mclip_set_option(name.encode(), int(value)), "mclip_set_option")
        self._ws_size.clear()          # plans (and therefore workspace sizes) may depend on the switches

    def get_option(self, name: str) -> int:
        v = ctypes.c_int(0)
        _check(self.lib, self.lib.mclip_get_option(name.encode(), ctypes.byref(v)), "mclip_get_option")
        return v.value

    def loss_finalize(self, row_lse, col_lse, diag, ls):
        dev = self._prep(row_lse, col_lse, diag, ls)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_loss_finalize(_ptr(row_lse), _ptr(col_lse), _ptr(diag), row_lse.numel(), _ptr(ls),
                                              _ptr(out), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_loss_finalize")
        return out

    def normalize_rows(self, x, out_dtype, eps, out=None):
        """`out` (optional): a [M, D] row-major view to write into -- e.g. the rank's own slot of an all-gather buffer."""
        dev = self._prep(x)
        M, D = x.shape
        if out is None:
            y = torch.empty((M, D), dtype=out_dtype, device=dev)
        else:
            if out.shape != (M, D) or out.dtype != out_dtype or out.stride(1) != 1 or out.device != dev:
                raise ValueError("normalize_rows: `out` must be a [M, D] row-major tensor of the output dtype on the same device")
            y = out
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_normalize_rows(_ptr(x), M, D, x.stride(0), float(eps), DTYPE_CODES[out_dtype], _ptr(y),
                                               y.stride(0), None, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_normalize_rows")
        return y

    def normalize_rows_bwd(self, x, g, eps):
        dev = self._prep(x, g)
        M, D = x.shape
        dx = torch.empty((M, D), dtype=torch.float32, device=dev)
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_normalize_rows_bwd(_ptr(x), _ptr(g), M, D, x.stride(0), g.stride(0), DTYPE_CODES[g.dtype],
                                                   float(eps), _ptr(dx), dx.stride(0),
                                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_normalize_rows_bwd")
        return dx

    def dls_finalize(self, u, v, diag, go, scale):
        """-> (t, dls); `v` / `diag` may be None (taken as zero)."""
        dev = self._prep(u, v, diag)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        t_out, dls_out = out[0:1], out[1:2]
        with self._DeviceGuard(dev):
            rc = self.lib.mclip_dls_finalize(_ptr(u), _ptr(v), _ptr(diag), u.numel(), _ptr(go), float(scale),
                                             _ptr(t_out), _ptr(dls_out),
                                             ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(self.lib, rc, "mclip_dls_finalize")
        return out[0], out[1]


_backend = None
_override = None


def set_backend_override(obj):
    """TEST-ONLY hook: route the primitive ops to `obj` (None restores the CUDA library).  The product has no CPU path;
    tests/ install an oracle-backed stand-in to run the multi-rank host logic under gloo on CPU.  Refused unless the
    process opted in with MCLIP_ALLOW_TEST_BACKEND=1 (tests/conftest.py sets it), so that no deployment can end up on a
    stand-in by accident."""
    global _override
    if obj is not None and os.environ.get("MCLIP_ALLOW_TEST_BACKEND") != "1":
        raise RuntimeError("set_backend_override is a test hook: set MCLIP_ALLOW_TEST_BACKEND=1 (tests/conftest.py does) to use it")
    _override = obj


def get_backend():
    global _backend
    if _override is not None:
        return _override
    if _backend is None:
        _backend = CudaBackend()
    return _backend
