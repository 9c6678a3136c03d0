#!/usr/bin/env python
"""Headline benchmark: contrastive loss fwd+bwd samples/s @ 32k global batch, D=512, bf16 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = ClipLoss forward + backward (local_loss=True, gather_with_grad=True) over the global batch,
rank r holding rows [r*B/N, (r+1)*B/N): config C3 of BASELINE.md, fixed global batch (strong scaling).
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU oracle port of the reference's ClipLoss
(the reference is pure Python/torch; /root/reference does not exist on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GLOBAL_BATCH = 32768
DIM = 512
LOGIT_SCALE = 14.2857  # 1/0.07, CLIP init
METRIC = "contrastive loss fwd+bwd samples/sec @32k global batch, D=512 bf16"
UNIT = "samples/s"
FALLBACK_PEAK_TFLOPS = 1590.0  # /opt/skills/guides/B200_PROFILING.md fallback (burst)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
        except Exception:
            pass
    return FALLBACK_PEAK_TFLOPS, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ----------------------------------------------------------------------------------------------------
def cpu_port_measure(sample_batch: int, iters: int, warmup: int):
    """Times the oracle port of the reference's W=1 ClipLoss (fp32 upcast of bf16 values, all host threads)
    on a bounded sample: `sample_batch` rows against `sample_batch` columns.  Per-sample cost of this path
    grows linearly with the batch, so samples/s at the 32k global batch = measured / (32768 / sample_batch)."""
    import torch
    from oracle import clip_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, txt = O.make_features(sample_batch, DIM, seed=1234, dtype=torch.bfloat16)
    img, txt = img.float(), txt.float()
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.ref_port_single(img, txt, LOGIT_SCALE)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    measured = sample_batch / t
    scaled = measured * sample_batch / GLOBAL_BATCH
    return {"value": scaled, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port of reference ClipLoss (loss.py:109-111,142-145), fp32, W=1, B={sample_batch} x D={DIM}, "
                      f"{len(times)} timed fwd+bwd of {t * 1e3:.1f} ms = {measured:.0f} samples/s at B={sample_batch}; "
                      f"scaled x{sample_batch}/{GLOBAL_BATCH} to the 32k batch (cost per sample is linear in B)"}, t


def workload_config(batch, world):
    return {"workload": "C3 large-batch contrastive loss: global B=%d, D=%d bf16, local_loss=True, gather_with_grad=True" % (batch, DIM),
            "global_batch": batch, "per_gpu_batch": batch // world, "dim": DIM, "logit_scale": LOGIT_SCALE}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 8192
    base, t = cpu_port_measure(sample_batch, iters=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * (GLOBAL_BATCH / sample_batch) ** 2,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(GLOBAL_BATCH, max(1, args.gpus)),
                       note="CPU arm (oracle port of the reference ClipLoss, fp32 upcast of the bf16 values, all host threads): each step "
                            "is a bounded B=8192 sample of the workload; value and ms_per_step are scaled to B=32768"),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH, help="global batch (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--max-seconds", type=float, default=900.0, help="hard wall-clock limit per process (watchdog)")
    args = ap.parse_args()
    watchdog = threading.Timer(args.max_seconds, lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from mamba_clip_b200 import ClipLoss, _cabi
    from oracle import clip_oracle as O  # input generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    B, D = args.batch, DIM
    assert B % world == 0
    Bl = B // world
    img_all, txt_all = O.make_features(B, D, seed=1234, dtype=torch.bfloat16)   # global problem, sliced by rank
    img_h = img_all[rank * Bl:(rank + 1) * Bl].contiguous().pin_memory()
    txt_h = txt_all[rank * Bl:(rank + 1) * Bl].contiguous().pin_memory()
    img_d = img_h.to(dev).requires_grad_(True)
    txt_d = txt_h.to(dev).requires_grad_(True)
    ls = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    import mamba_clip_b200
    mamba_clip_b200.enable_cuda_graphs(not args.no_graphs)   # public switch: replay the captured launch sequences
    crit = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    be = _cabi.get_backend()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(a, b):
        a.grad = b.grad = ls.grad = None
        loss = crit(image_features=a, text_features=b, logit_scale=ls)["contrastive_loss"]
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ----
    c0 = be.launch_count()
    step(img_d, txt_d)                     # always eager (graphs are captured on the third call of a shape)
    launches_per_step = be.launch_count() - c0
    for _ in range(max(args.warmup - 1, 3 if not args.no_graphs else 0)):
        step(img_d, txt_d)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = be.launch_count()
    be.kernel_timing(True)                 # bracket every launch of the dominant kernel inside the timed region with CUDA events
    barrier()
    t_wall0 = time.perf_counter()
    for s, e in evs:
        flush.zero_()                      # evict inputs from L2 between timed iterations (untimed)
        s.record()
        step(img_d, txt_d)
        e.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    insitu_ms, insitu_n = be.kernel_timing(False)   # (0, 0) when the step replays CUDA graphs: events cannot be recorded there
    launches = be.launch_count() - n0
    if not args.no_graphs:
        # replayed graphs do not pass through the library's launch counter: same kernels as the eager step counted above
        launches = launches_per_step * args.steps
    step_ms = [s.elapsed_time(e) for s, e in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    value = B / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers (H2D of the step's inputs + D2H of the loss) ----
    # Every step's features start in pinned host memory and its loss ends in pinned host memory.  The copies of
    # step k+1 are issued on a copy stream while step k computes (two device buffer sets), the way a training loop
    # with non_blocking=True transfers behaves; all copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(img_d).requires_grad_(True), torch.empty_like(txt_d).requires_grad_(True)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_h = torch.empty(args.steps + 3, dtype=torch.float32).pin_memory()

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last used this slot is done with it
            with torch.no_grad():
                bufs[slot][0].copy_(img_h, non_blocking=True)
                bufs[slot][1].copy_(txt_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(n):
        main = torch.cuda.current_stream(dev)
        for sl in range(2):
            consumed[sl].record(main)
        stage_inputs(0)
        for k in range(n):
            slot = k & 1
            if k + 1 < n:
                stage_inputs(slot ^ 1)
            main.wait_event(ready[slot])
            loss = step(*bufs[slot])
            consumed[slot].record(main)
            loss_h[k].copy_(loss.detach(), non_blocking=True)

    e2e_run(3)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_run(args.steps)
    s1.record()
    barrier()
    e2e_ms = torch.tensor([s0.elapsed_time(s1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = B / (float(e2e_ms) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel (block_grad, one side) timed alone on this rank's shapes ----
    with torch.no_grad():
        all_t = txt_all.to(dev) if world > 1 else txt_d.detach()
        lsv = torch.full((1,), LOGIT_SCALE, device=dev)
        go = torch.ones(1, device=dev)
        row_lse, _ = be.row_lse(img_d.detach(), all_t, lsv, rank * Bl, False)
        col_lse, _ = be.row_lse(all_t, img_d.detach(), lsv, -rank * Bl, False)  # statistic only: shape-correct LSE
        for _ in range(3):
            be.block_grad(img_d.detach(), all_t, lsv, go, row_lse, col_lse, rank * Bl, 1.0, 1.0, 2.0, 0.5 / Bl)
        torch.cuda.synchronize(dev)
        kt = []
        for _ in range(10):
            flush.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            be.block_grad(img_d.detach(), all_t, lsv, go, row_lse, col_lse, rank * Bl, 1.0, 1.0, 2.0, 0.5 / Bl)
            k1.record()
            torch.cuda.synchronize(dev)
            kt.append(k0.elapsed_time(k1))
        k_iso_ms = sum(kt) / len(kt)
    # the roofline uses the launches of the timed region itself when they could be bracketed (eager steps), else the
    # isolated launches above
    k_ms = insitu_ms / insitu_n if insitu_n > 0 else k_iso_ms
    peak, peak_sustained, peak_src = load_peaks()
    alg_flops_launch = 2.0 * Bl * B * D            # dX = G @ Y: one of the three algorithmic GEMMs (S recompute not counted)
    achieved = alg_flops_launch / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and B == GLOBAL_BATCH:
        try:
            traffic = json.load(open(tpath)).get("block_grad_dram_bytes_per_launch")
        except Exception:
            traffic = None
    step_alg_tflops = 6.0 * B * B * D / world / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "tc_block_grad2_kernel (S recompute + dX = G@Y for one side; launched twice per step)",
                "kernel_ms": k_ms, "kernel_ms_source": (f"CUDA events around each of the {insitu_n} launches inside the timed region" if insitu_n > 0
                                                           else "10 isolated launches after the timed region (the step replays CUDA graphs)"),
                "kernel_ms_isolated": k_iso_ms, "peak_source": f"{peak_src} burst bf16 (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback",
                "executed_tflops": achieved * 2.0,
                "step_algorithmic_tflops_per_gpu": step_alg_tflops, "step_frac_of_peak": step_alg_tflops / peak,
                "step_frac_of_sustained_peak": step_alg_tflops / peak_sustained}

    if rank == 0:
        cpu_base = None
        if not args.no_cpu_baseline:
            cpu_base, _ = cpu_port_measure(8192, iters=5, warmup=1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(B, world), **{
                       "parallelism": f"dp{world} (row/column blocks per rank, NCCL all-gather of features + LSE vectors)",
                       "cuda_graphs": not args.no_graphs,
                       "cuda_graphs_replayed": any(len(g.graphs) > 0 for g in mamba_clip_b200._function._graph_cache.values()),
                       "l2": "256 MiB flush between timed iterations", "timing": "CUDA events per step, max over ranks",
                       "wall_s_timed_region": t_wall}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * Bl * D * 2, "d2h_bytes_per_step": 4,
                    "ms_per_step": float(e2e_ms), "last_loss": float(loss_h[args.steps - 1]),
                    "note": "inputs from pinned host memory every step (H2D prefetched on a copy stream, double-buffered), loss copied back to pinned host memory every step"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
