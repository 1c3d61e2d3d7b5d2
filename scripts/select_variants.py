#!/usr/bin/env python
"""GPU-side A/B of libtcmp<suffix>.so builds on the goal-IK selection kernel (K7): one subprocess per library (each
under its own timeout: a variant with CTA barriers must not be able to hang the box), results hashed so that every
variant is checked to return the default build's output bit for bit.  Kernel tuning aid, not the bench."""
import glob, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from torque_constrained_motion_planning_b200 import _lib
    _lib.LIB_PATH = sys.argv[1]
    from torque_constrained_motion_planning_b200 import engine
    from bench import Q_LO, Q_HI
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    out = {}
    h = hashlib.sha256()
    for n, nf in ((200_000, 25), (200_003, 32), (1001, 70), (37, 5)):
        q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
        free = np.empty((nf, n)); free[0] = q[6]; free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
        qd, fd = torch.as_tensor(q, device=dev), torch.as_tensor(free, device=dev)
        trans, rot = engine.fk_batch(qd)

        def t(fn, reps=5):
            for _ in range(2): fn()
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
        for mode in ("base", "rne", "dyn"):
            res = engine.ik_select(rot, trans, fd, qd, 3.0, mode=mode)
            for a in res:
                h.update(np.ascontiguousarray(a.cpu().numpy() if hasattr(a, "cpu") else a).tobytes())
            if n >= 200_000:
                ms = t(lambda: engine.ik_select(rot, trans, fd, qd, 3.0, mode=mode))
                out["%s_nf%d_Gsolves" % (mode, nf)] = round(n * nf / ms / 1e6, 3)
    out["sha"] = h.hexdigest()[:16]
    print(json.dumps(out))
    sys.exit(0)

for path in sorted(glob.glob(os.path.join(ROOT, "torque_constrained_motion_planning_b200", "libtcmp*.so"))):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), path], capture_output=True, text=True, timeout=90)
    except subprocess.TimeoutExpired:
        print(os.path.basename(path), "TIMEOUT"); continue
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(os.path.basename(path), "FAILED", r.stderr[-600:]); continue
    print("%-24s" % os.path.basename(path), line[-1], flush=True)
