"""bench.py's reference arm runs on CPU: check the JSON line contract (keys, types, the arm's own e2e / cpu_baseline)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "cpu_baseline", "e2e"}


def run_reference(extra_env=None, args=()):
    env = dict(os.environ, **(extra_env or {}))
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--poses-per-gpu", "65536",
                        "--no-extra", *args], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run_reference()
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "collision-checked poses/s" and d["unit"] == "poses/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # same `config` object as the GPU arm prints for the same flags (the driver compares the two arms' configs)
    assert d["config"]["poses_per_gpu_per_step"] == 65536 and d["config"]["parallelism"] == "single GPU"
    assert d["ms_per_step"] * 1e-3 * d["value"] == __import__("pytest").approx(65536, rel=1e-6)
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb


def test_reference_arm_uses_all_cores_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still use every core it may run on"""
    d = run_reference({"OMP_NUM_THREADS": "1"})
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun (N > 1) rank 0 alone runs and prints; the other ranks exit 0 without work"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_json_line_on_the_gpu():
    """a short run of the GPU arm: every key of the contract is present and consistent"""
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--poses-per-gpu", "1048576", "--no-extra",
                        "--cpu-sample", "65536"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (KEYS - {"impl"}) | {"roofline", "clocks", "gpu_launches"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] == 3 and d["dtype"] == "f32" and d["scaling"] == "weak"
    r = d["roofline"]
    # the committed ncu counters are for the 2^24-pose launch: at this batch size the line falls back to the HBM view
    assert r["bound"] in ("hbm", "issue") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    h = r["hbm"]
    assert h["unit"] == "GB/s" and abs(h["achieved"] - 25 * 1048576 / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * h["achieved"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 24 * 1048576 and e["d2h_bytes_per_step"] == 1048576 and 0 < e["value"] < d["value"]
    assert e["h2d_probe"]["gbs_aggregate"] > 1 and 0 < e["frac_of_h2d_probe"] < 1.5
    assert e["edges"]["edges_per_s"] > 0 and e["edges"]["h2d_bytes_per_step"] == 96 * (1 << 20)
    assert d["config"]["poses_per_gpu_per_step"] == 1048576
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
