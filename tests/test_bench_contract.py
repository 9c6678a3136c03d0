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


def _load_bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_keeps_the_samples_of_the_timed_window():
    m = _load_bench()
    idle = "1965, 1965, 120.5, Not Active, Not Active, Not Active, Not Active"
    rows = [(0.10, idle), (1.00, "1700, 1965, 640.0, Not Active, Not Active, Not Active, Active"),
            (1.02, "1650, 1965, 650.0, Not Active, Not Active, Not Active, Active"), (1.30, idle), (1.40, "garbage")]
    r = m.ClockSampler.summarise(rows, 0.99, 1.03)
    assert r["samples"] == 2 and r["sm_mhz"] == 1675.0 and r["reasons"] == ["sw_power_cap"] and r["power_w_max"] == 650.0
    r = m.ClockSampler.summarise(rows, 0.50, 0.51)                 # window shorter than a period: nearest samples, and says so
    assert r["samples"] == 2 and "nearest" in r["window"]
    assert m.ClockSampler.summarise([], 0.0, 1.0)["samples"] == 0


def test_clock_sampler_process_is_started_filtered_and_terminated(tmp_path, monkeypatch):
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text('#!/bin/bash\nwhile true; do echo "1800, 1965, 500.0, Not Active, Not Active, Not Active, Active"; sleep 0.02; done\n')
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    m = _load_bench()
    c = m.ClockSampler(0, max_seconds=10)
    c.start()
    time.sleep(0.3)
    t0 = time.perf_counter()
    time.sleep(0.1)
    r = c.stop(t0, time.perf_counter())
    assert 2 <= r["samples"] <= 10 and r["sm_mhz"] == 1800.0 and r["reasons"] == ["sw_power_cap"]
    assert c.proc.poll() is not None            # no sampler left behind
