#!/usr/bin/env python
"""GPU-side A/B timing of experimental libtcmp<suffix>.so builds (kernel tuning aid, not the bench)."""
import ctypes, glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import sample_states
from torque_constrained_motion_planning_b200 import _lib

N = int(os.environ.get("N", 1_000_000))
sets = []
for s in range(4):
    sets.append(tuple(torch.as_tensor(a, device="cuda") for a in sample_states(N, 2 + s)))
tau = torch.empty((7, N), dtype=torch.float64, device="cuda")
mask = torch.empty((N,), dtype=torch.uint8, device="cuda")
ref = None
for path in sorted(glob.glob(os.path.join(ROOT, "torque_constrained_motion_planning_b200", "libtcmp*.so"))):
    lib = ctypes.CDLL(path)
    f = lib.tcmp_rne_batch
    f.restype, f.argtypes = _lib.SIGNATURES["tcmp_rne_batch"]
    res = {}
    for label, mode, want_tau in [("rne", 0, True), ("rne_mask", 0, False), ("nov", 1, True), ("dyn", 2, True)]:
        def run(i):
            q, qd, qdd, m = sets[i % 4]
            rc = f(mode, 0, N, q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), m.data_ptr(), 0.0, 0.01,
                   tau.data_ptr() if want_tau else None, mask.data_ptr(), None)
            assert rc == 0
        for i in range(5): run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        e0.record(torch.cuda.default_stream())
        for i in range(K): run(i)
        e1.record(torch.cuda.default_stream())
        torch.cuda.synchronize()
        res[label] = N * K / (e0.elapsed_time(e1) * 1e-3) / 1e9
    run(0)
    q, qd, qdd, m = sets[0]
    f(0, 0, N, q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), m.data_ptr(), 0.0, 0.01, tau.data_ptr(), mask.data_ptr(), None)
    torch.cuda.synchronize()
    if ref is None:
        ref = tau.clone()
    print("%-28s" % os.path.basename(path), " ".join("%s=%.2fG/s" % kv for kv in res.items()),
          "maxdiff_vs_first=%.2e" % (tau - ref).abs().max().item(), flush=True)
