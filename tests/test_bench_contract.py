"""bench.py's CPU-runnable leg: `--impl reference` (the host OpenMP arm) must print exactly ONE JSON line with the
keys the measurement contract names, also under a torchrun-style environment where only rank 0 may speak."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spmv_gflops" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm prints the SAME config object our arm prints for this workload, and says what it timed
    assert d["config"] == bench.workload_config("cfg1", None, 1, "fused", False) and d["dtype"] == "f64"
    assert d["config"]["rows"] == 1_000_000 and d["config"]["nnz"] == 4_996_000
    assert d["steps"] == 2 and d["warmup"] == 3 and "2 timed calls" in cb["sample"]


def test_reference_arm_is_silent_on_other_ranks():
    assert run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_clock_sampler_counts_only_lines_after_the_mark_and_waits_for_one(tmp_path, monkeypatch):
    """bench.ClockSampler against a fake nvidia-smi that needs a second before its first line (a fresh box): lines
    printed before mark() do not count; if none has arrived by stop(), the GPU is kept busy until one does; throttle
    reasons are collected; the child is gone afterwards"""
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\n"
                    "echo '0, 1500, 1965, 300, 0x0, Not Active, Not Active, Not Active, Not Active'\n"
                    "sleep 1.2\n"
                    "while true; do echo '0, 1950, 1965, 900, 0x4, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.3)                   # the first (idle, 1500 MHz) line is in; the steady ones are a second away
    s.mark()
    busy = []
    out = s.stop(keep_busy=lambda: (busy.append(1), time.sleep(0.05)))
    assert len(busy) >= 5             # waited ~0.9 s for the first line after the mark, with the GPU kept busy
    assert out["samples"] >= 1 and out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert s.proc.poll() is not None


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop(keep_busy=lambda: None)
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]
