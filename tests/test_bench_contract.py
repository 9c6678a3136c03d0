"""bench.py's command-line contract where it can be exercised without a GPU: the reference arm (`--impl reference`, the
reference's own CPU ClipLoss from oracle/_ref, or the oracle port where the copy is absent) prints exactly one JSON line
with the keys the driver reads, alone and under torchrun (rank 0 only); the GPU arm refuses to run without a GPU (no CPU
fallback).  Small sizes (config C2: per-rank batch 64) so this stays in seconds."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LINE_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _json_lines(stdout):
    return [json.loads(l) for l in stdout.splitlines() if l.startswith("{")]


def _check_reference_line(d, n_gpus, batch):
    assert LINE_KEYS <= set(d), sorted(LINE_KEYS - set(d))
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["gpu_launches"] == 0
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["global_batch"] == batch and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] - batch / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--config", "C2", "--steps", "5", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1, 64)
    assert lines[0]["steps"] == 5 and lines[0]["warmup"] == 3


def test_reference_arm_under_torchrun_only_rank0_reports():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(_free_port()), "bench.py", "--impl", "reference", "--gpus", "2",
                        "--config", "C2", "--steps", "3", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1                      # the other rank exits 0 without work
    _check_reference_line(lines[0], 2, 128)     # the CPU arm times W=1 over the same GLOBAL batch (2 x 64)


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, "bench.py", "--config", "C2", "--steps", "2", "--warmup", "3", "--max-seconds", "120"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert _json_lines(r.stdout) == []          # no number from a CPU path
