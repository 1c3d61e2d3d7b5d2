#!/usr/bin/env python
"""CPU soak of K1's arithmetic (host build of csrc/panda_model.cuh: customised recursion + table-driven sincos) against
the C oracle: N_CHUNKS x 1 M config-2 states, every fourth chunk with half of its joint values snapped to multiples of
pi/4 or to the joint limits.  Prints one JSON line.  No GPU needed."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from conftest import Q_HI, Q_LO, sample_states

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 20
so = os.path.join(ROOT, "tests", "native", "librne_host.so")
if not os.path.exists(so):
    raise SystemExit("run pytest tests/test_rne_core_host.py once to build " + so)
L = ctypes.CDLL(so)
dp = ctypes.POINTER(ctypes.c_double)
p = lambda a: np.ascontiguousarray(a).ctypes.data_as(dp)
worst, mism, total, margin = 0.0, 0, 0, np.inf
lim = np.array([87.0, 87, 87, 87, 12, 12])[:, None]
t0 = time.time()
for chunk in range(n_chunks):
    q, qd, qdd, mass = sample_states(1_000_000, seed=500 + chunk)
    if chunk % 4 == 3:
        rng = np.random.default_rng(900 + chunk)
        snap = rng.random(q.shape) < 0.5
        q = np.where(snap, np.clip(np.round(q / (np.pi / 4)) * (np.pi / 4), Q_LO[:, None], Q_HI[:, None]), q)
    n = q.shape[1]
    tau, ok = np.empty((7, n)), np.empty(n, np.uint8)
    L.host_rne_batch_table(ctypes.c_int64(n), p(q), p(qd), p(qdd), p(mass), ctypes.c_double(0.01), p(tau),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    tau_o, ok_o = oracle.torque_test_batch("rne", q, qd, qdd, mass)
    worst = max(worst, float(np.abs(tau - tau_o).max()))
    mism += int((ok != ok_o).sum())
    margin = min(margin, float(np.abs(lim - np.abs(tau_o[:6])).min()))
    total += n
print(json.dumps({"what": "host build of rne_core_table (K1's arithmetic) vs C oracle, rne mode, config-2 states",
                  "states": total, "mask_mismatches": mism, "max_abs_torque_error_Nm": worst,
                  "smallest_distance_of_a_torque_to_its_limit_Nm": margin, "seconds": round(time.time() - t0, 1)}))
