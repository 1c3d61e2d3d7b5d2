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


def test_round2_bench_evidence_follows_the_contract():
    """profiles/r02: the final build's driver-command lines.  Beyond the round-1 keys: `launch` (one CUDA graph), `settle`,
    `sustained` as a first-class key, `roofline.traffic` read from the committed ncu CSV, the executed FP64 count, the
    planner extras with a MEASURED reference time, the host-link ceiling, and a reference arm within 10 % of the GPU arm's
    cpu_baseline on the same box."""
    prof = os.path.join(ROOT, "profiles", "r02")
    d = json.loads(open(os.path.join(prof, "bench_n1_final.json")).read())
    ref = json.loads(open(os.path.join(prof, "bench_reference_n1_final.json")).read())
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] >= 3 and d["gpu_launches"] == 20
    assert d["launch"].startswith("one CUDA graph") and d["settle"]["seconds"] > 0
    assert abs(d["value"] - 1e6 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-6
    assert abs(d["sustained"]["value"] / d["value"] - 1) < 0.05          # same clock regime: K = 20 agrees with 1.5 s
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic_source"].startswith("profiles/r02/")
    assert 150e6 < r["traffic"] <= 233e6 and r["executed"]["fp64_instr_per_state"] <= 600
    assert 0.5 < r["executed"]["pipe_frac"] < r["frac"] <= 1.0
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["config"]["workload"] == d["config"]["workload"]
    assert abs(ref["value"] / d["cpu_baseline"]["value"] - 1) < 0.10 and ref["cpu_baseline"]["cores"] == d["cpu_baseline"]["cores"]
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 176_000_000 and e["d2h_bytes_per_step"] == 57_000_000
    assert 0.8 < e["link"]["in_call_share_of_ceiling"] <= 1.2 and e["pipelined"]["value"] > e["value"]
    for p in d["extras"]["planner"]:
        assert p["gpu_strict_trajectory_equals_reference"] is True and p["reference_measured_s"] > 20
        assert p["gpu_strict_s"] < 0.2 and p["gpu_batched_s"] < p["gpu_strict_s"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])
    values = {}
    for n in (1, 2, 4, 8):
        s = json.loads([l for l in open(os.path.join(prof, "scale_n%d.json" % n)) if l.startswith("{")][-1])
        assert s["n_gpus"] == n and s["scaling"] == "weak" and s["launch"].startswith("one CUDA graph")
        assert s["gpu_launches"] == (20 if n == 1 else 60)            # N > 1: scatter kernel + signal + wait per step
        if n > 1:
            assert s["gather_check"].startswith("peer-store gather") and "tcmp_peer_signal" in s["gather"]
        assert 0.8 < s["e2e"]["link"]["in_call_share_of_ceiling"] <= 1.2
        values[n] = s["value"]
    assert values[8] / (8 * values[1]) > 0.90 and values[2] / (2 * values[1]) > 0.95
