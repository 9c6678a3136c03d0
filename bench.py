#!/usr/bin/env python
"""Headline benchmark: contrastive loss fwd+bwd samples/s @ 32k global batch, D=512, bf16 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C3|C4|C2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = ClipLoss forward + backward over the global batch, rank r holding rows [r*B/N, (r+1)*B/N).
  C3 (default, the BASELINE metric): global B=32768, D=512, bf16, local_loss=True, gather_with_grad=True   (strong scaling)
  C4: global B=65536, D=768, bf16, same mode (ViT-L/14-shaped; the no-materialised-logits proof)           (strong scaling)
  C2: per-rank B=64, D=512, bf16, local_loss=False, gather_with_grad=False (the repo's DDP config)        (weak scaling, latency)
Before anything is timed, every rank checks loss / dI / dT / d(logit_scale) of one step against the chunked fp64 oracle
(loss and d ls over all rows, gradients on a seeded row subset) and the line carries the result as "parity"; above the
2e-3 bar of bf16 inputs the run exits non-zero.  Prints ONE JSON line on rank 0.
`--impl reference` times the reference's own CPU ClipLoss (oracle/_ref = its unmodified loss.py when present, else the
oracle port) on the host cores at the configuration's true size.
"""
from __future__ import annotations

import argparse
import atexit
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

LOGIT_SCALE = 14.2857  # 1/0.07, CLIP init
UNIT = "samples/s"
FALLBACK_PEAK_TFLOPS = 1590.0  # /opt/skills/guides/B200_PROFILING.md fallback (burst)
PARITY_BAR = 2e-3              # bf16 inputs, fp32 accumulation (BASELINE.json north_star)

CONFIGS = {
    "C3": dict(batch=32768, dim=512, local_loss=True, gwg=True, scaling="strong",
               metric="contrastive loss fwd+bwd samples/sec @32k global batch, D=512 bf16",
               workload="C3 large-batch contrastive loss: global B=%(B)d, D=%(D)d bf16, local_loss=True, gather_with_grad=True"),
    "C4": dict(batch=65536, dim=768, local_loss=True, gwg=True, scaling="strong",
               metric="contrastive loss fwd+bwd samples/sec @64k global batch, D=768 bf16",
               workload="C4 ViT-L/14-shaped contrastive loss: global B=%(B)d, D=%(D)d bf16, local_loss=True, gather_with_grad=True"),
    "C2": dict(per_rank=64, dim=512, local_loss=False, gwg=False, scaling="weak",
               metric="contrastive loss fwd+bwd samples/sec @64 per-rank batch, D=512 bf16 (latency config)",
               workload="C2 stage-1 DDP config: per-rank B=64 (global B=%(B)d), D=%(D)d bf16, local_loss=False, gather_with_grad=False"),
}


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
    """nvidia-smi clocks / throttle reasons, one sample every 20 ms, each stamped with its arrival time; `stop(t0, t1)` keeps
    the samples that arrived inside [t0, t1] (time.perf_counter(): the device-timed region through the end of the e2e
    region).  If none did (a region shorter than one period), the samples closest to the window are used and the line says so."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, max_seconds: int = 960):
        self.index = index
        self.max_seconds = int(max_seconds)
        self.rows = []          # (arrival time, csv line)
        self.proc = None
        self.thread = None

    def start(self):
        try:
            # bounded lifetime even if this process dies without cleaning up (the watchdog uses os._exit)
            self.proc = subprocess.Popen(["timeout", "-s", "TERM", str(self.max_seconds), "nvidia-smi", "-i", str(self.index),
                                          f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            atexit.register(self._kill)
        except Exception:
            self.proc = None
            return

        def pump():
            try:
                for line in self.proc.stdout:
                    self.rows.append((time.perf_counter(), line.strip()))
            except Exception:
                pass
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def _kill(self):
        try:
            if self.proc is not None and self.proc.poll() is None:
                self.proc.terminate()
        except Exception:
            pass

    @classmethod
    def summarise(cls, rows, t0=None, t1=None):
        parsed = []
        for t, r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                parsed.append((t, float(f[0]), float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[3:7]]))
            except ValueError:
                continue
        window = "all samples"
        if t0 is not None and t1 is not None and parsed:
            inside = [p for p in parsed if t0 <= p[0] <= t1 + 0.03]       # a sample lands up to one period after it was taken
            if inside:
                parsed, window = inside, "samples that arrived inside the timed regions (device-timed steps through the e2e steps)"
            else:
                mid = 0.5 * (t0 + t1)
                parsed = sorted(parsed, key=lambda p: abs(p[0] - mid))[:2]
                window = "timed regions shorter than one 20 ms sampling period: the two samples nearest to them"
        sm = [p[1] for p in parsed]
        mx = [p[2] for p in parsed]
        pw = [p[3] for p in parsed]
        reasons = sorted({n for p in parsed for n, on in zip(cls.NAMES, p[4]) if on})
        busy = [s_ for s_, p_ in zip(sm, pw) if p_ > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None, "window": window}

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        try:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            return self.summarise(list(self.rows), t0, t1)
        except Exception as e:          # never let the clock report take the bench line down
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling failed: %r" % (e,)], "samples": 0}


def resolve_config(args, world):
    c = dict(CONFIGS[args.config])
    if "per_rank" in c:
        c["B"] = (args.batch // world if args.batch else c["per_rank"]) * world
    else:
        c["B"] = args.batch or c["batch"]
    c["D"] = args.dim or c["dim"]
    if args.batch or args.dim:   # a non-BASELINE size: say so in the metric
        c["metric"] = "contrastive loss fwd+bwd samples/sec @%d global batch, D=%d bf16 (non-BASELINE size)" % (c["B"], c["D"])
    c["workload"] = c["workload"] % c
    return c


def workload_config(cfg, world):
    return {"workload": cfg["workload"], "global_batch": cfg["B"], "per_gpu_batch": cfg["B"] // world, "dim": cfg["D"],
            "logit_scale": LOGIT_SCALE, "local_loss": cfg["local_loss"], "gather_with_grad": cfg["gwg"]}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ----------------------------------------------------------------------------------------------------
def reference_cpu_measure(batch: int, dim: int, iters: int, warmup: int):
    """Times the reference's own ClipLoss at W=1 on the host cores (fp32 upcast of the bf16 values, every host thread):
    oracle/_ref (the unmodified loss.py) when it is present, else the oracle's line-for-line port.
    -> (samples/s, seconds per fwd+bwd, kind, cores, description)."""
    import torch
    from oracle import build_ref, clip_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, txt = O.make_features(batch, dim, seed=1234, dtype=torch.bfloat16)
    img, txt = img.float(), txt.float()
    if build_ref.available():
        crit = build_ref.load().ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1)
        kind, what = "reference", "unmodified reference ClipLoss (oracle/_ref copy of src/mamba_clip/loss.py)"

        def one():
            a = img.clone().requires_grad_(True)
            b = txt.clone().requires_grad_(True)
            s = torch.tensor(LOGIT_SCALE, requires_grad=True)
            crit(image_features=a, text_features=b, logit_scale=s)["contrastive_loss"].backward()
    else:
        kind, what = "port", "oracle port of the reference ClipLoss (loss.py:109-111,142-145)"

        def one():
            O.ref_port_single(img, txt, LOGIT_SCALE)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return batch / t, t, kind, cores, what


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    cfg = resolve_config(args, world)
    B, D = cfg["B"], cfg["D"]
    note = None
    if B * B * 4 * 5 > 100e9:
        # the reference materialises two [B, B] fp32 logits blocks plus their log-softmax / gradient temporaries
        # (~5 B^2 floats): C4 does not fit host memory -- time the largest size that does
        note = "B=%d needs ~%.0f GB of logits on the CPU: timed at B=32768 and scaled by B/32768 (cost per sample is linear in B)" % (
            B, B * B * 4 * 5 / 1e9)
        Bt = 32768
    else:
        Bt = B
    if Bt >= 8192:
        iters, warm = max(1, min(args.steps, 3)), 1     # ~4 s per fwd+bwd at B=32768 on the box's 16 cores; BASELINE.md section 4: 1 + 3
    else:
        iters, warm = max(1, args.steps), max(1, args.warmup)   # milliseconds per step: the flags as given
    value, t, kind, cores, what = reference_cpu_measure(Bt, D, iters=iters, warmup=warm)
    if Bt != B:
        value = value * Bt / B
        t = t * (B / Bt) ** 2
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{what}, fp32, W=1, B={Bt} x D={D}: {warm} warm-up + {iters} timed fwd+bwd" + (f"; {note}" if note else "")}
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": iters, "warmup": warm, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(cfg, world),
                       note="CPU arm: the reference's ClipLoss at W=1 over the same global batch (fp32 upcast of the bf16 values, all host threads)"
                            + ("; " + note if note else "")),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned staging buffers are allocated,
    so that first touch puts them on the GPU's NUMA node (8 ranks pulling from one node's memory halve each other's
    H2D bandwidth).  Best effort: silently skipped when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def parity_check(step, cfg, img_all, txt_all, a, b, ls, rank, world, dev, n_rows=256):
    """One step of the product path against the chunked fp64 oracle.  Rank 0 computes the global statistics (two fp64
    passes over S, never materialised), every rank then checks its own loss, d(logit_scale) (both exact over all of
    its rows / columns) and dI, dT on `n_rows` seeded rows of its shard.  -> dict of relative errors."""
    import torch
    import torch.distributed as dist
    from oracle import clip_oracle as O
    B = cfg["B"]
    Bl = B // world
    t0 = time.perf_counter()
    loss = step(a, b)
    torch.cuda.synchronize(dev)
    got = (float(loss.detach()), a.grad.detach().float().cpu(), b.grad.detach().float().cpu(), float(ls.grad))
    vec = torch.empty(5, B, dtype=torch.float64)
    if rank == 0:
        torch.set_num_threads(os.cpu_count() or 1)
        st = O.global_stats(img_all.float(), txt_all.float(), LOGIT_SCALE)
        vec = torch.stack((st.row_lse, st.col_lse, st.u, st.v, st.diag))
    if world > 1:
        vd = vec.to(dev)
        dist.broadcast(vd, src=0)
        vec = vd.cpu()
    st = O.GlobalStats(vec[0], vec[1], vec[2], vec[3], vec[4])
    g = torch.Generator().manual_seed(4321 + rank)
    rows = torch.randperm(Bl, generator=g)[:min(n_rows, Bl)].sort().values
    r_loss, r_di, r_dt, r_dls = O.closed_form_rows(img_all.float(), txt_all.float(), LOGIT_SCALE, world, rank,
                                                   cfg["local_loss"], cfg["gwg"], st, rows)
    err = torch.tensor([abs(got[0] - float(r_loss)) / abs(float(r_loss)), O.rel_err(got[1][rows], r_di),
                        O.rel_err(got[2][rows], r_dt), abs(got[3] - float(r_dls)) / abs(float(r_dls))], dtype=torch.float64)
    if world > 1:
        ed = err.to(dev)
        dist.all_reduce(ed, op=dist.ReduceOp.MAX)
        err = ed.cpu()
    return {"loss": float(err[0]), "dI": float(err[1]), "dT": float(err[2]), "dls": float(err[3]),
            "rows": int(rows.numel()), "ranks": world, "bar": PARITY_BAR,
            "oracle": "chunked fp64 closed form (oracle.clip_oracle.global_stats + closed_form_rows) on the fp32 upcast of the same bf16 "
                      "values: loss and d(logit_scale) over all rows/columns of every rank, dI/dT on %d seeded rows per rank "
                      "against all columns; max over ranks" % int(rows.numel()),
            "seconds": time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the global batch of the chosen config")
    ap.add_argument("--dim", type=int, default=0, help="override the feature dimension of the chosen config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing oracle check (development only)")
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None

    import torch
    import torch.distributed as dist
    from mamba_clip_b200 import ClipLoss, _cabi
    from oracle import clip_oracle as O  # input generator, parity check and cpu_baseline leg only

    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    cfg = resolve_config(args, world)
    B, D = cfg["B"], cfg["D"]
    assert B % world == 0
    Bl = B // world
    img_all, txt_all = O.make_features(B, D, seed=1234, dtype=torch.bfloat16)   # global problem, sliced by rank
    # one pinned staging buffer per rank: [image shard; text shard] -> ONE H2D copy per step in the e2e leg
    stage_h = torch.empty((2, Bl, D), dtype=torch.bfloat16).pin_memory()
    stage_h[0].copy_(img_all[rank * Bl:(rank + 1) * Bl])
    stage_h[1].copy_(txt_all[rank * Bl:(rank + 1) * Bl])
    img_d = stage_h[0].to(dev).requires_grad_(True)
    txt_d = stage_h[1].to(dev).requires_grad_(True)
    ls = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    import mamba_clip_b200
    mamba_clip_b200.enable_cuda_graphs(not args.no_graphs)   # public switch: replay the captured launch sequences
    crit = ClipLoss(local_loss=cfg["local_loss"], gather_with_grad=cfg["gwg"], cache_labels=True, rank=rank, world_size=world)
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

    # nvidia-smi needs 0.1-0.3 s before its first sample and the timed region of a multi-GPU run lasts tens of milliseconds:
    # start it now (it streams a sample every 20 ms from here on) and keep only the samples that arrive inside the timed
    # regions (ClockSampler.stop).
    sampler = ClockSampler(local_rank, max_seconds=int(args.max_seconds) + 30)
    if rank == 0:
        sampler.start()

    # ---- parity of the exact path that is timed below (first call: eager; graphs replay the same kernels) ----
    c0 = be.launch_count()
    step(img_d, txt_d)                     # always eager (graphs are captured on the third call of a shape)
    launches_per_step = be.launch_count() - c0
    parity = None
    if not args.no_parity:
        parity = parity_check(step, cfg, img_all, txt_all, img_d, txt_d, ls, rank, world, dev)
        worst = max(parity[k] for k in ("loss", "dI", "dT", "dls"))
        if not (worst <= PARITY_BAR):
            if rank == 0:
                print(json.dumps({"error": "parity check failed", "parity": parity}), flush=True)
            raise SystemExit(4)

    # ---- device-resident timing ----
    for _ in range(max(args.warmup - 1, 3 if not args.no_graphs else 0)):
        step(img_d, txt_d)
    barrier()
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
    graphs_replayed = any(len(g.graphs) > 0 for g in mamba_clip_b200._function._graph_cache.values())
    if graphs_replayed:
        # replayed graphs do not pass through the library's launch counter: same kernels as the eager step counted above
        launches = launches_per_step * args.steps
    step_ms = [s.elapsed_time(e) for s, e in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    value = B / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers (H2D of the step's inputs + D2H of the loss) ----
    # Every step's features start in pinned host memory and its loss ends in pinned host memory.  The copy of step k+1 is
    # issued on a copy stream while step k computes (two device buffer sets), the way a training loop with
    # non_blocking=True transfers behaves; all copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty((2, Bl, D), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    views = [(b[0].requires_grad_(True), b[1].requires_grad_(True)) for b in bufs]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_h = torch.empty(args.steps + 3, dtype=torch.float32).pin_memory()

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last used this slot is done with it
            with torch.no_grad():
                bufs[slot].copy_(stage_h, non_blocking=True)
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
            loss = step(*views[slot])
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
    barrier()
    clocks = sampler.stop(t_wall0, time.perf_counter()) if rank == 0 else None

    # ---- dominant kernel(s) of the step timed alone on this rank's shapes ----
    # W = 1, D <= 512: the shared-recompute backward (mclip_fused_grad: tc_block_grad2_kernel<store G> panels on the main
    # stream overlapped with tc_gemm_tn_kernel on a side stream) -- one bracket around the whole call, algorithmic flops
    # = dI + dT = 4 B^2 D, executed 6 B^2 D.  Otherwise: one launch of the one-sided recompute kernel (dX = G Y,
    # algorithmic 2 B_l B D, executed 4 B_l B D), two of which make a backward.
    fused = world == 1 and be.fused_supported(img_d.detach(), txt_d.detach())
    with torch.no_grad():
        all_t = txt_all.to(dev) if world > 1 else txt_d.detach()
        lsv = torch.full((1,), LOGIT_SCALE, device=dev)
        go = torch.ones(1, device=dev)
        row_lse, _ = be.row_lse(img_d.detach(), all_t, lsv, rank * Bl, False)
        col_lse, _ = be.row_lse(all_t, img_d.detach(), lsv, -rank * Bl, False)  # statistic only: shape-correct LSE

        def dominant():
            if fused:
                be.fused_grad(img_d.detach(), all_t, lsv, go, row_lse, col_lse, 0, 0.5 / Bl)
            else:
                be.block_grad(img_d.detach(), all_t, lsv, go, row_lse, col_lse, rank * Bl, 1.0, 1.0, 2.0, 0.5 / Bl)
        for _ in range(3):
            dominant()
        torch.cuda.synchronize(dev)
        kt = []
        for _ in range(10):
            flush.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            dominant()
            k1.record()
            torch.cuda.synchronize(dev)
            kt.append(k0.elapsed_time(k1))
        k_iso_ms = sum(kt) / len(kt)
    # the roofline uses the launches of the timed region itself when they could be bracketed (eager steps), else the
    # isolated launches above
    k_ms = insitu_ms / insitu_n if insitu_n > 0 else k_iso_ms
    peak, peak_sustained, peak_src = load_peaks()
    alg_units, exe_units = (2.0, 3.0) if fused else (1.0, 2.0)
    alg_flops_launch = alg_units * 2.0 * Bl * B * D     # S recompute is never counted
    achieved = alg_flops_launch / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and args.config == "C3" and not (args.batch or args.dim):
        try:
            traffic = json.load(open(tpath)).get("fused_grad_dram_bytes_per_call" if fused else "block_grad_dram_bytes_per_launch")
        except Exception:
            traffic = None
    step_alg_tflops = 6.0 * B * B * D / world / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": ("mclip_fused_grad: tc_block_grad2_kernel<store G> (S recompute + dI = G@T per row panel) overlapped with "
                           "tc_gemm_tn_kernel (dT = G^T@I) -- the whole backward, one bracket per step" if fused else
                           "backward recompute kernel behind mclip_block_grad (S recompute + dX = G@Y for one side; launched twice per step)"),
                "algorithmic_gemm_units_per_bracket": alg_units, "executed_gemm_units_per_bracket": exe_units,
                "kernel_ms": k_ms, "kernel_ms_source": (f"CUDA events around each of the {insitu_n} launches inside the timed region" if insitu_n > 0
                                                           else "10 isolated launches after the timed region (the step replays CUDA graphs)"),
                "kernel_ms_isolated": k_iso_ms, "peak_source": f"{peak_src} burst bf16 (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback",
                "executed_tflops": achieved * exe_units / alg_units,
                "step_algorithmic_tflops_per_gpu": step_alg_tflops, "step_frac_of_peak": step_alg_tflops / peak,
                "step_frac_of_sustained_peak": step_alg_tflops / peak_sustained}

    if rank == 0:
        cpu_base = None
        if not args.no_cpu_baseline and world == 1:
            # bounded sample of the same workload: the reference's ClipLoss on B=8192 rows x 8192 columns of the same
            # seeded inputs (the --impl reference arm times the full size)
            sb = min(8192, B)
            v, t, kind, cores, what = reference_cpu_measure(sb, D, iters=5, warmup=1)
            cpu_base = {"value": v * sb / B, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{what}, fp32, W=1, B={sb} x D={D}, 5 timed fwd+bwd of {t * 1e3:.1f} ms = {v:.0f} samples/s at B={sb}; "
                                  f"scaled x{sb}/{B} to the global batch (cost per sample is linear in B; --impl reference times the full size)"}
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(cfg, world), **{
                       "parallelism": f"dp{world} (row/column blocks per rank, NCCL all-gather of features + statistics)",
                       "cuda_graphs": not args.no_graphs, "cuda_graphs_replayed": graphs_replayed,
                       "l2": "256 MiB flush between timed iterations", "timing": "CUDA events per step, max over ranks",
                       "wall_s_timed_region": t_wall, "numa_bound_cpus": numa_cpus}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * Bl * D * 2, "d2h_bytes_per_step": 4,
                    "ms_per_step": float(e2e_ms), "last_loss": float(loss_h[args.steps - 1]),
                    "note": "inputs from pinned host memory every step (one packed H2D per step, prefetched on a copy stream, double-buffered), loss copied back to pinned host memory every step"},
            "gpu_launches": launches,
            "us_per_step": ms_per_step * 1e3,
            "parity": parity,
            "roofline": roofline,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
