#!/usr/bin/env python
"""Developer diagnostics for the tcgen05 kernels (not part of the test-suite): compares the tensor-core
path with the FFMA path and the CPU oracle on a few shapes, prints error statistics and timings.
Each case runs in its own process with a timeout so that a hung kernel cannot take the box down."""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(kind, M, N, D, ls, diag_off, dtype, oracle):
    import torch
    from mamba_clip_b200 import _cabi
    from oracle import clip_oracle as O
    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    g = torch.Generator().manual_seed(M + 3 * N)
    x = torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=-1)
    y = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    k = min(M, N)
    y[:k] = torch.nn.functional.normalize(x[:k] + 0.1 * torch.randn(k, D, generator=g), dim=-1)
    x, y = x.to(dt), y.to(dt)
    xd, yd = x.cuda(), y.cuda()
    lsd = torch.tensor([ls], device="cuda")
    tc, simt = _cabi.CudaBackend(path=2), _cabi.CudaBackend(path=1)
    if kind == "fwd":
        a, da = tc.row_lse(xd, yd, lsd, diag_off, True)
        torch.cuda.synchronize()
        b, db = simt.row_lse(xd, yd, lsd, diag_off, True)
        torch.cuda.synchronize()
        print(f"fwd {M}x{N}x{D} ls={ls}: tc-vs-simt lse max|d|={float((a - b).abs().max()):.3e} diag max|d|={float((da - db).abs().max()):.3e}"
              f" nan={int(torch.isnan(a).sum())}")
        if float((a - b).abs().max()) > 1e-2 or torch.isnan(a).any():
            bad = ((a - b).abs() > 1e-2) | torch.isnan(a)
            idx = bad.nonzero().flatten()[:16].tolist()
            print("  first bad rows:", idx, "tc:", a[idx].tolist(), "simt:", b[idx].tolist())
            print("  bad count:", int(bad.sum()), "of", M)
        if oracle:
            r, rd = O.block_row_lse(x.float(), y.float(), ls, diag_off)
            print(f"  tc-vs-oracle lse max|d|={float((a.cpu().double() - r).abs().max()):.3e}")
    else:
        lx, _ = simt.row_lse(xd, yd, lsd, 0, False)
        ly, _ = simt.row_lse(yd, xd, lsd, 0, False)
        go = torch.tensor([2.0], device="cuda")
        want_rd = os.environ.get("MCLIP_BWD_V3") != "1"
        a, ra = tc.block_grad(xd, yd, lsd, go, lx, ly, diag_off, 1.0, 1.0, 2.0, 0.5 / M, want_rd)
        torch.cuda.synchronize()
        b, rb = simt.block_grad(xd, yd, lsd, go, lx, ly, diag_off, 1.0, 1.0, 2.0, 0.5 / M)
        torch.cuda.synchronize()
        af, bf = a.float(), b.float()
        rel = float((af - bf).norm() / bf.norm().clamp_min(1e-30))
        rdd = float((ra - rb).abs().max()) if ra is not None else float("nan")
        print(f"bwd {M}x{N}x{D} ls={ls}: tc-vs-simt dX rel={rel:.3e} |simt|={float(bf.norm()):.3e} rowdot max|d|={rdd:.3e}"
              f" nan={int(torch.isnan(af).sum())}")
        if oracle:
            ref, _ = O.block_grad(x.float(), y.float(), ls, lx.cpu(), ly.cpu(), diag_off, 1.0, 1.0, 2.0, 2.0 * ls * 0.5 / M)
            print(f"  vs fp64 oracle: tc rel={O.rel_err(a.cpu(), ref):.3e}  simt rel={O.rel_err(b.cpu(), ref):.3e}  (kernel: {'1-CTA' if os.environ.get('MCLIP_BWD_1CTA') == '1' else ('pair v2' if want_rd else 'pair v3 transposed')})")
        if rel > 5e-3 or torch.isnan(af).any():
            err = (af - bf).abs()
            rows = err.max(dim=1).values
            cols = err.max(dim=0).values
            print("  worst rows:", rows.topk(min(8, M)).indices.tolist(), "worst cols:", cols.topk(min(8, D)).indices.tolist())
            print("  per-64-col-block err:", [round(float(err[:, c:c + 64].max()), 5) for c in range(0, D, 64)])
            print("  per-32-row-block err:", [round(float(err[r:r + 32].max()), 5) for r in range(0, min(M, 256), 32)])
            print("  sample tc :", af[0, :8].tolist())
            print("  sample ref:", bf[0, :8].tolist())


def time_case(kind, M, N, D, dtype, iters):
    import torch
    from mamba_clip_b200 import _cabi
    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    x = torch.nn.functional.normalize(torch.randn(M, D, device="cuda"), dim=-1).to(dt)
    y = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=-1).to(dt)
    lsd = torch.tensor([14.2857], device="cuda")
    tc = _cabi.CudaBackend(path=2)
    lx, _ = tc.row_lse(x, y, lsd, 0, False)
    ly, _ = tc.row_lse(y, x, lsd, 0, False)
    go = torch.tensor([1.0], device="cuda")
    fn = (lambda: tc.row_lse(x, y, lsd, 0, True)) if kind == "fwd" else (
        lambda: tc.block_grad(x, y, lsd, go, lx, ly, 0, 1.0, 1.0, 2.0, 0.5 / M, os.environ.get("MCLIP_BWD_V3") != "1"))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * M * N * D * (1 if kind == "fwd" else 2)
    print(f"time {kind} {M}x{N}x{D}: {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s useful ({'S' if kind == 'fwd' else 'S+dX'})")


CASES = [
    ("fwd", 128, 256, 64, 10.0, 0), ("fwd", 128, 256, 512, 14.2857, 0), ("fwd", 128, 512, 512, 14.2857, 0),
    ("fwd", 129, 300, 512, 100.0, 64), ("fwd", 512, 4096, 512, 30.0, 1024), ("fwd", 300, 1000, 768, 30.0, 17),
    ("bwd", 128, 128, 64, 10.0, 0), ("bwd", 128, 128, 256, 10.0, 0), ("bwd", 128, 128, 512, 14.2857, 0),
    ("bwd", 128, 256, 512, 14.2857, 0), ("bwd", 129, 300, 512, 30.0, 64), ("bwd", 512, 4096, 512, 30.0, 1024),
    ("bwd", 300, 1000, 768, 30.0, 17), ("bwd", 64, 64, 64, 14.2857, 0), ("bwd", 1000, 3000, 384, 14.2857, 100), ("bwd", 2048, 2048, 512, 14.2857, 0),
    ("bwd_v3", 512, 4096, 512, 30.0, 1024), ("bwd_v3", 1000, 3000, 384, 14.2857, 100), ("bwd_2exp", 1000, 3000, 384, 14.2857, 100),
    ("fwd", 256, 256, 512, 14.2857, 0), ("fwd", 2048, 2048, 512, 100.0, 0), ("fwd", 1000, 3000, 384, 14.2857, 100),
    ("bwd", 4096, 4096, 512, 14.2857, 0), ("bwd", 640, 1111, 200, 14.2857, 300),
]
TIMES = [("fwd", 8192, 8192, 512), ("bwd", 8192, 8192, 512), ("fwd", 32768, 32768, 512), ("bwd", 32768, 32768, 512)]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs="*")
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--dbg-sweep", action="store_true")
    args = ap.parse_args()
    if args.one:
        a = args.one
        if a[0] in ("time", "time_1cta"):
            time_case(a[1], int(a[2]), int(a[3]), int(a[4]), "bf16", 5)
        else:
            run_case(a[0], int(a[1]), int(a[2]), int(a[3]), float(a[4]), int(a[5]), "bf16", int(a[1]) * int(a[2]) <= 1 << 22)
        sys.exit(0)
    if args.dbg_sweep:
        for kind, extra in (("bwd", {}), ("bwd", {"MCLIP_BWD_V3": "1"}), ("bwd", {"MCLIP_DBG": "8"}), ("bwd", {"MCLIP_DBG": "16"})):
            env = dict(os.environ, **extra)
            for shape in (("32768", "32768", "512"), ("4096", "32768", "512")):
                r = subprocess.run([sys.executable, __file__, "--one", "time", kind, *shape], capture_output=True,
                                   text=True, timeout=150, env=env)
                print(f"{extra}:", r.stdout.strip(), r.stderr[-300:] if r.returncode else "", flush=True)
        sys.exit(0)
    jobs = [[c[0]] + [str(v) for v in c[1:]] for c in CASES]
    if not args.no_time:
        jobs += [["time"] + [str(v) for v in t] for t in TIMES]
        jobs += [["time_1cta", "bwd", "32768", "32768", "512"]]
    for j in jobs:
        t0 = time.time()
        env = dict(os.environ)
        if j[0].endswith("_1cta"):
            j = [j[0][:-5]] + j[1:]
            env["MCLIP_BWD_1CTA"] = "1"
        if j[0].endswith("_v3"):
            j = [j[0][:-3]] + j[1:]
            env["MCLIP_BWD_V3"] = "1"
        if j[0].endswith("_2exp"):
            j = [j[0][:-5]] + j[1:]
            env["MCLIP_DBG"] = "8"
        try:
            r = subprocess.run([sys.executable, __file__, "--one"] + j, capture_output=True, text=True, timeout=150, env=env)
            out = (r.stdout + ("\n[stderr] " + r.stderr[-1500:] if r.returncode != 0 else "")).strip()
            print(out, f"[rc={r.returncode} {time.time() - t0:.1f}s]", flush=True)
        except subprocess.TimeoutExpired:
            print("TIMEOUT (hang?)", j, flush=True)
            break
