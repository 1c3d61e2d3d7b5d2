#!/usr/bin/env python
"""GPU-side A/B of libtcmp<suffix>.so builds on the non-K1 workloads (IK sweep, edges, goal-IK selection, model
kernel): one subprocess per library, bench.run_extras + two more timings.  Kernel tuning aid, not the bench."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from torque_constrained_motion_planning_b200 import _lib
    _lib.LIB_PATH = sys.argv[1]
    from torque_constrained_motion_planning_b200 import engine
    import bench
    dev = torch.device("cuda:0")
    ex = bench.run_extras(engine, dev, 10)
    out = {"ik_counts_G": ex["ik"]["solves_per_s_counts_only"] / 1e9, "ik_sols_G": ex["ik"]["solves_per_s_with_solutions"] / 1e9,
           "fk_G": ex["ik"]["fk_poses_per_s"] / 1e9, "edges_G": ex["edges"]["edges_per_s"] / 1e9}
    q, qd, qdd, m = (torch.as_tensor(a, device=dev) for a in bench.sample_states(1_000_000, 2))
    tau = torch.empty((7, 1_000_000), dtype=torch.float64, device=dev)
    ok = torch.empty(1_000_000, dtype=torch.uint8, device=dev)
    mdl = engine.InertialModel.default()

    def timeit(fn, reps):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps
    out["model_rne_G"] = 1e6 / timeit(lambda: engine.torque_test_batch(q, qd, qdd, m, out_tau=tau, out_mask=ok, model=mdl), 30) / 1e9
    # goal-IK selection: 200k poses x 25 free values
    rng = np.random.default_rng(5)
    n = 200_000
    qq = rng.uniform(bench.Q_LO[:, None], bench.Q_HI[:, None], size=(7, n))
    tr, ro = engine.fk_batch(torch.as_tensor(qq, device=dev))
    free = torch.as_tensor(rng.uniform(-2.8973, 2.8973, size=(25, n)), device=dev)
    qref = torch.as_tensor(qq[:, :1].copy().reshape(7), device=dev)
    t = timeit(lambda: engine.ik_select(ro, tr, free, qref, payload_mass=1.0), 5)
    out["select_Gsolves"] = n * 25 / t / 1e9
    print(json.dumps(out))
    sys.exit(0)

for path in sorted(glob.glob(os.path.join(ROOT, "torque_constrained_motion_planning_b200", "libtcmp*.so"))):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), path], capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(os.path.basename(path), "FAILED", r.stderr[-600:]); continue
    d = json.loads(line[-1])
    print("%-24s" % os.path.basename(path), " ".join("%s=%.3f" % kv for kv in d.items()), flush=True)
