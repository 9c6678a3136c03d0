#!/usr/bin/env python
"""Per-primitive timings at the headline shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mamba_clip_b200 import _cabi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ls_val = float(sys.argv[3]) if len(sys.argv) > 3 else 14.2857
be = _cabi.CudaBackend(path=2)
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=-1).bfloat16()
y = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=-1).bfloat16()
ls = torch.tensor([ls_val], device="cuda")
go = torch.ones(1, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(name, fn, n=8, flops=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    extra = f"  {flops / ms / 1e9:8.1f} TFLOP/s" if flops else ""
    print(f"{name:34s} {ms:8.4f} ms (min {min(ts):.4f}){extra}", flush=True)
    return ms


unit = 2.0 * B * B * D
diag, ref, status = be.pair_ref(x, y, ls, 0)
timeit("pair_ref (diag + reference)", lambda: be.pair_ref(x, y, ls, 0))
timeit("pair_lse (two-sided fwd)", lambda: be.pair_lse(x, y, ls, ref, status, True), flops=unit)
print("status", int(status.item()))
row_lse, rowdot, col_lse = be.pair_lse(x, y, ls, ref, status, True)
timeit("row_lse (one-sided fwd)", lambda: be.row_lse(x, y, ls, 0, True, True), flops=unit)
timeit("row_lse predicated off", lambda: be.row_lse(x, y, ls, 0, False, True, run_if=status, out_lse=row_lse, out_rowdot=rowdot))
r1 = be.row_lse(x, y, ls, 0, True, True)
r2 = be.row_lse(y, x, ls, 0, False, True)
print("max |row_lse pair - one-sided|", float((row_lse - r1[0]).abs().max()), " col:", float((col_lse - r2[0]).abs().max()),
      " rowdot:", float((rowdot - r1[2]).abs().max()))
timeit("block_grad (one side)", lambda: be.block_grad(x, y, ls, go, r1[0], r2[0], 0, 1.0, 1.0, 2.0, 0.5 / B, False), flops=unit)
timeit("block_grad (one side, +rowdot)", lambda: be.block_grad(x, y, ls, go, r1[0], r2[0], 0, 1.0, 1.0, 2.0, 0.5 / B, True), flops=unit)
