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
