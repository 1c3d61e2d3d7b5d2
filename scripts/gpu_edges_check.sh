#!/bin/bash
# Edge-kernel parity tests + one timing of configs[3] (100k edges x 64 waypoints, rne, 5 kg).
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_guards.py tests/test_gpu_dropin.py -x -q -k "edge or traj or planner or graph" 2>&1 | tail -4
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
from torque_constrained_motion_planning_b200 import engine
qa, qb = (torch.as_tensor(a, device="cuda") for a in bench.sample_edges(100_000, 4))
for mode in ("rne", "dyn", "nov"):
    for _ in range(3): engine.edge_feasibility(qa, qb, 64, 5.0, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): engine.edge_feasibility(qa, qb, 64, 5.0, mode=mode)
    e1.record(); torch.cuda.synchronize()
    print(mode, "%.4f G edges/s" % (20 * 100_000 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
PY
