#!/bin/bash
# Round-end evidence on one box: GPU test suite, smoke, the driver's N = 1 bench command for both arms, then (if the
# box has them) the 1/2/4/8 weak-scaling lines.  gpurun [--gpus 8] -- bash scripts/gpu_final.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_n1_final.json 2> gpurun_out/bench_reference_n1_final.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_reference_n1_final.json").read()); d=json.loads(open("gpurun_out/bench_n1_final.json").read())
print("reference arm %.3e states/s on %d threads; cpu_baseline %.3e" % (r["value"], r["cpu_baseline"]["cores"], d["cpu_baseline"]["value"]))
print("gpu value %.4e ms/step %.5f sustained %.4e e2e %.4e -> e2e ratio %.1f, device ratio %.0f" % (d["value"], d["ms_per_step"], d["sustained"]["value"], d["e2e"]["value"], d["e2e"]["value"]/r["value"], d["value"]/r["value"]))
print("roofline", {k:d["roofline"][k] for k in ("achieved","peak","frac","peak_sustained","frac_of_sustained_peak","kernel_ms","traffic")}, d["roofline"]["executed"], d["roofline"]["hbm"]["frac"])
print("modes", d["modes"]); print("ik", d["extras"]["ik"]["solves_per_s_counts_only"], d["extras"]["ik"]["solves_per_s_with_solutions"], "edges", d["extras"]["edges"]["edges_per_s"])
for p in d["extras"]["planner"]: print("planner", p["scene"][:11], p["gpu_strict_s"], p["gpu_batched_s"], p["gpu_batched_arrays_s"], p["reference_measured_s"], p["gpu_strict_trajectory_equals_reference"])
print("link", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["e2e"]["link"].items() if k not in ("note","affinity")}); print("clocks", d["clocks"])
print("cpu other", d["cpu_baseline"].get("other_workloads"))
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 8 ]; then bash scripts/gpu_scale.sh 1 2 4 8; fi
