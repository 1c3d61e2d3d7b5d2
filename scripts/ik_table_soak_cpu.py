#!/usr/bin/env python
"""CPU soak of the IK core built with TCMP_IK_TABLE_SINCOS=1 (host build of csrc/ik_core.cuh) against the compiled,
unmodified reference solver: solution counts over N_CHUNKS x 2 M solves (random reachable poses x 25 free values).
Prints one JSON line.  No GPU needed."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from test_ik_core_host import _build_host_ik
from conftest import Q_HI, Q_LO

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 25
VARIANT = os.environ.get("IK_VARIANT", "table")
L = _build_host_ik("libik_host_table.so", ["TCMP_IK_TABLE_SINCOS=1"]) if VARIANT == "table" else _build_host_ik("libik_host.so")
dp = ctypes.POINTER(ctypes.c_double)
total = mism = 0
hist = np.zeros(9, dtype=np.int64)
t0 = time.time()
for chunk in range(n_chunks):
    rng = np.random.default_rng(1000 + chunk)
    n, nf = 80_000, 25
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    if chunk % 5 == 4:      # every fifth chunk: joint values snapped to multiples of pi/4 (singular branches)
        snap = rng.random(q.shape) < 0.3
        q = np.where(snap, np.clip(np.round(q / (np.pi / 4)) * (np.pi / 4), Q_LO[:, None], Q_HI[:, None]), q)
    trans, rot = oracle.ref_fk_batch(q)
    free = np.ascontiguousarray(np.vstack([q[6:7], rng.uniform(Q_LO[6], Q_HI[6], size=(nf - 1, n))]))
    _, want = oracle.ref_ik_batch(rot, trans, free, want_sols=False)
    got = np.zeros(n * nf, np.int32)
    L.host_ik_batch(ctypes.c_int64(n), rot.ctypes.data_as(dp), trans.ctypes.data_as(dp), free.ctypes.data_as(dp), nf, 0,
                    None, got.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None)
    total += n * nf
    mism += int((got != want).sum())
    hist += np.bincount(want, minlength=9)[:9]
print(json.dumps({"variant": ("TCMP_IK_TABLE_SINCOS=1" if VARIANT == "table" else "default (libm sincos)") + " (host build of csrc/ik_core.cuh)", "solves": total,
                  "count_mismatches_vs_compiled_reference": mism, "count_histogram": hist.tolist(),
                  "seconds": round(time.time() - t0, 1)}))
