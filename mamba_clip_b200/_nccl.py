"""Direct NCCL calls (ctypes on the libnccl that torch already loaded) for the per-step collectives of the loss.

torch.distributed stays the plumbing: it bootstraps the communicator (the NCCL unique id travels through a c10d
broadcast) and serves every collective that is not on the per-step path.  The per-step all-gathers, however, are
latency-bound (a 4-32 MB feature gather and two O(B) statistic gathers per step), and a c10d collective costs
~100 us of host time per call (work objects, watchdog registration, stream hand-offs) -- more than the NCCL kernel
itself runs.  `ncclAllGather` issued straight onto the compute stream costs a few microseconds of host time, needs no
cross-stream events and can be captured into the CUDA graph of the step together with the kernels around it.

MCLIP_DIRECT_NCCL=0 (or any failure to load / initialise) falls back to c10d collectives.
"""
from __future__ import annotations

import atexit
import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

_NCCL_UNIQUE_ID_BYTES = 128
_NCCL_UINT8 = 1      # ncclUint8 / ncclChar family: the gathers are type-agnostic byte moves


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * _NCCL_UNIQUE_ID_BYTES)]


_lib = None
_comms = {}


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL("libnccl.so.2")     # same soname torch links: resolves to the already-loaded library
        lib.ncclGetUniqueId.restype = ctypes.c_int
        lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
        lib.ncclCommInitRank.restype = ctypes.c_int
        lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        lib.ncclAllGather.restype = ctypes.c_int
        lib.ncclAllGather.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]
        lib.ncclGroupStart.restype = ctypes.c_int
        lib.ncclGroupStart.argtypes = []
        lib.ncclGroupEnd.restype = ctypes.c_int
        lib.ncclGroupEnd.argtypes = []
        lib.ncclCommDestroy.restype = ctypes.c_int
        lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        lib.ncclCommAbort.restype = ctypes.c_int
        lib.ncclCommAbort.argtypes = [ctypes.c_void_p]
        lib.ncclGetErrorString.restype = ctypes.c_char_p
        lib.ncclGetErrorString.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


class DirectComm:
    """One NCCL communicator over the ranks of a torch.distributed group, driven without c10d."""

    def __init__(self, group, device: torch.device):
        lib = _load()
        self.lib = lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        uid = _UniqueId()
        if self.rank == 0:
            self._check(lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        t = torch.tensor(list(bytes(uid.internal)), dtype=torch.uint8, device=device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        raw = bytes(t.cpu().tolist())
        ctypes.memmove(ctypes.byref(uid), raw, _NCCL_UNIQUE_ID_BYTES)
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            self._check(lib.ncclCommInitRank(ctypes.byref(comm), self.world, uid, self.rank), "ncclCommInitRank")
        self.comm = comm
        self.device = device

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.lib.ncclGetErrorString(rc).decode()}")

    def all_gather(self, out: torch.Tensor, x: torch.Tensor) -> None:
        """out[rank * n : (rank + 1) * n] <- x of every rank, on the current CUDA stream (in order, no events)."""
        assert out.is_contiguous() and x.is_contiguous() and out.numel() * out.element_size() == self.world * x.numel() * x.element_size()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.ncclAllGather(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                           x.numel() * x.element_size(), _NCCL_UINT8, self.comm, ctypes.c_void_p(stream)),
                    "ncclAllGather")


    def all_gather_many(self, pairs) -> None:
        """Several all-gathers as ONE grouped NCCL operation (one launch) on the current CUDA stream."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.ncclGroupStart(), "ncclGroupStart")
        try:
            for out, x in pairs:
                assert out.is_contiguous() and x.is_contiguous() and out.numel() * out.element_size() == self.world * x.numel() * x.element_size()
                self._check(self.lib.ncclAllGather(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                   x.numel() * x.element_size(), _NCCL_UINT8, self.comm, ctypes.c_void_p(stream)),
                            "ncclAllGather")
        finally:
            self._check(self.lib.ncclGroupEnd(), "ncclGroupEnd")


_DIRECT = os.environ.get("MCLIP_DIRECT_NCCL", "1") != "0"     # read once at import, not per step


def direct_comm(group, device: torch.device) -> Optional[DirectComm]:
    """The (cached) direct communicator for `group` on `device`, or None when direct NCCL is disabled / unavailable.
    Creation is a collective: every rank of the group reaches it in its first multi-rank forward.  There is exactly ONE
    direct communicator per (group, device) and every call on it is issued on the caller's current stream, in program
    order -- collectives of different communicators are never in flight concurrently on a device.
    The cache entry holds a reference to the group object, so its id cannot be reused while the entry lives."""
    if not _DIRECT or device.type != "cuda":
        return None
    if dist.get_backend(group) != "nccl":
        return None
    pg = group if group is not None else dist.group.WORLD
    key = (id(pg), device.index)
    hit = _comms.get(key)
    if hit is not None and hit[0] is pg:
        return hit[1]
    try:
        comm = DirectComm(group, device)
    except (OSError, AttributeError) as e:   # library not loadable: every rank fails the same way
        import warnings
        warnings.warn(f"mamba_clip_b200: direct NCCL unavailable ({e}); using torch.distributed collectives")
        comm = None
    _comms[key] = (pg, comm)
    return comm


def destroy_all(abort: bool = False) -> None:
    """Release every direct communicator.  Call it (collectively) before `destroy_process_group` in long-lived
    processes that re-create groups.  At interpreter exit the registered handler uses ncclCommAbort instead, which
    never waits for peers that may already be gone."""
    for _, c in list(_comms.values()):
        if c is not None:
            try:
                (c.lib.ncclCommAbort if abort else c.lib.ncclCommDestroy)(c.comm)
            except Exception:
                pass
    _comms.clear()


atexit.register(destroy_all, True)
