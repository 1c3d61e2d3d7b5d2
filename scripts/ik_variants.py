#!/usr/bin/env python
"""GPU-side A/B of experimental IK builds (libtcmp<suffix>.so next to the product library; kernel tuning aid, not the
bench).  For every library: solves/s of tcmp_ik_batch on configs[2] (1M reachable poses x 25 free values; counts only,
and a 200k-pose slice with the [8][7] solution sets), and the number of solution-count mismatches against the
compiled reference on the structured singular-pose families (tests/ik_families.py).

    python scripts/ik_variants.py build      # CPU container: cross-compile the variants
    python scripts/ik_variants.py            # GPU box
"""
import ctypes
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = os.path.join(ROOT, "torque_constrained_motion_planning_b200")

VARIANTS = {
    "_ik_table": dict(defines=["TCMP_IK_TABLE_SINCOS=1"]),
    "_ik_mb8": dict(defines=["TCMP_IK_MINBLOCKS=8"]),           # 128 registers: 16 warps / SM instead of 12
    "_ik_mb7": dict(defines=["TCMP_IK_MINBLOCKS=7"]),
    "_ik_table_mb8": dict(defines=["TCMP_IK_TABLE_SINCOS=1", "TCMP_IK_MINBLOCKS=8"]),
}


def build_variants():
    from torque_constrained_motion_planning_b200 import build as tb
    for suffix, kw in VARIANTS.items():
        print(tb.build(force=True, suffix=suffix, **kw))


def main():
    import torch
    import oracle
    from ik_families import structured_families
    from torque_constrained_motion_planning_b200 import _lib
    Q_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
    Q_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
    NT = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(3)
    n, nf = 1_000_000, 25
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    free = np.empty((nf, n))
    free[0] = q[6]
    free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
    trans, rot = oracle.ref_fk_batch(q)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    d_rot, d_trans, d_free = dev(rot), dev(trans), dev(free)
    counts = torch.empty(n * nf, dtype=torch.int32, device="cuda")
    ns = 200_000
    s_rot, s_trans, s_free = dev(rot[:, :ns]), dev(trans[:, :ns]), dev(free[:, :ns])
    sols = torch.empty((ns * nf, 8, 7), dtype=torch.float64, device="cuda")
    fams = []
    for name, (fq, ffree) in structured_families(n_per=8000, seed=11).items():
        ft, fr = oracle.ref_fk_batch(fq)
        _, cr = oracle.ref_ik_batch(fr, ft, ffree, want_sols=False, nthreads=NT)
        fams.append((name, dev(fr), dev(ft), dev(ffree), cr))
    _, cr_rand = oracle.ref_ik_batch(rot[:, :ns], trans[:, :ns], free[:, :ns], want_sols=False, nthreads=NT)

    for path in sorted(glob.glob(os.path.join(PKG, "libtcmp*.so"))):
        lib = ctypes.CDLL(path)
        f = lib.tcmp_ik_batch
        f.restype, f.argtypes = _lib.SIGNATURES["tcmp_ik_batch"]

        def run(r, t, fr, so, co, nn):
            rc = f(nn, r.data_ptr(), t.data_ptr(), fr.data_ptr(), fr.shape[0], 0, None if so is None else so.data_ptr(),
                   co.data_ptr(), None, None)
            assert rc == 0

        def timeit(fn, reps):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.default_stream())
            for _ in range(reps):
                fn()
            e1.record(torch.cuda.default_stream())
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / reps

        t_cnt = timeit(lambda: run(d_rot, d_trans, d_free, None, counts, n), 5)
        t_sol = timeit(lambda: run(s_rot, s_trans, s_free, sols, counts, ns), 5)
        run(s_rot, s_trans, s_free, None, counts, ns)
        torch.cuda.synchronize()
        mism_rand = int((counts[:ns * nf].cpu().numpy() != cr_rand).sum())
        mism = {}
        for name, fr, ft, ffree, cr in fams:
            c = torch.empty(len(cr), dtype=torch.int32, device="cuda")
            run(fr, ft, ffree, None, c, fr.shape[1])
            torch.cuda.synchronize()
            m = int((c.cpu().numpy() != cr).sum())
            if m:
                mism[name] = m
        print(json.dumps({"lib": os.path.basename(path), "Gsolves_per_s_counts": n * nf / t_cnt / 1e9,
                          "Gsolves_per_s_sets": ns * nf / t_sol / 1e9, "mismatch_random_5M": mism_rand,
                          "structured_solves": int(sum(len(x[4]) for x in fams)),
                          "structured_mismatches": mism}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build_variants()
    else:
        main()
