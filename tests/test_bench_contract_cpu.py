"""CPU: the parts of bench.py's contract that need no GPU -- the `--impl reference` arm (the CPU implementation of the
path timed on the host cores) prints one JSON line with the driver's keys, for the metric BASELINE.json names."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert baseline["metric"].startswith(d["metric"])
    assert d["unit"] == "states/s" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_env_runs_on_rank_zero_only():
    """Launched as the driver launches N > 1 (RANK / WORLD_SIZE in the environment): rank 0 prints the line, every
    other rank exits 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29555")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_committed_bench_evidence_follows_the_contract():
    """profiles/r01/bench_n1_final.json and scale_n{2,4,8}.json are lines bench.py printed on the GPU box: they must
    carry the driver's keys plus `roofline`, `cpu_baseline` (N = 1), `e2e`, `gpu_launches` and a clean clock record."""
    prof = os.path.join(ROOT, "profiles", "r01")
    for name, n in (("bench_n1_final.json", 1), ("scale_n2.json", 2), ("scale_n4.json", 4), ("scale_n8.json", 8)):
        d = json.loads(open(os.path.join(prof, name)).read())
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
            assert key in d, (name, key)
        assert d["n_gpus"] == n and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["gpu_launches"] == d["steps"] > 0
        r = d["roofline"]
        assert r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert not bad & set(d["clocks"]["reasons"]) and d["clocks"]["sm_mhz"] > 0.8 * d["clocks"]["sm_max_mhz"]
        if n == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < d["value"]
            assert abs(d["value"] - 1e6 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-6
    ref = json.loads(open(os.path.join(prof, "bench_reference_n1_final.json")).read())
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["config"]["workload"] == d["config"]["workload"]
