"""Structured IK pose families that drive the analytic solver into its singular branches.

Used three ways:
  * tests (`tests/test_ik_core_host.py`, `tests/test_gpu_parity.py`) import `structured_families` for the count
    parity sweeps against the compiled reference (oracle/_ref);
  * `python scripts/ik_structured_families.py --compare` runs the host build of csrc/ik_core.cuh against the
    compiled reference over all families and prints the mismatches per family;
  * `python scripts/ik_structured_families.py --coverage /tmp/cov/libcov.so` drives a gcov-instrumented build of
    the reference (built out of tree, see profiles/r02/ik_reference_coverage.md) so the reached branches of the
    generated solver can be listed.

Every family is a set of joint vectors q[7][n] (pose = reference ComputeFk(q)) plus free-joint rows: the pose's own
joint 7 first, then special values and random values -- the shape of the reference's sweep (ikfast.py:153-159).
"""
import argparse
import ctypes
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ik_families import J4_SING, Q_HI, Q_LO, SPECIAL, structured_families, wrist_axis_family  # noqa: E402,F401


def _ref_lib(path):
    L = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    L.ref_ik_batch.argtypes = [ctypes.c_int64, dp, dp, dp, ctypes.c_int, ctypes.c_int, dp,
                               ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
    L.ref_ik_batch.restype = None
    return L


def main():
    import oracle
    ap = argparse.ArgumentParser()
    ap.add_argument("--coverage", default=None, help="path of a gcov-instrumented reference driver .so")
    ap.add_argument("--compare", action="store_true")
    ap.add_argument("--n-per", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=11)
    a = ap.parse_args()
    fams = structured_families(a.n_per, a.seed)
    host = None
    if a.compare:
        import test_ik_core_host as t
        host = t._wrap_host_ik(t._build_host_ik("libik_host.so"))
    cov = _ref_lib(a.coverage) if a.coverage else None
    dp = ctypes.POINTER(ctypes.c_double)
    total = bad = 0
    for name, (q, free) in fams.items():
        trans, rot = oracle.ref_fk_batch(q)
        n, nf = q.shape[1], free.shape[0]
        if cov is not None:
            c = np.zeros(n * nf, np.int32)
            rot_c, trans_c, free_c = (np.ascontiguousarray(x) for x in (rot, trans, free))
            cov.ref_ik_batch(n, rot_c.ctypes.data_as(dp), trans_c.ctypes.data_as(dp), free_c.ctypes.data_as(dp),
                             nf, 0, None, c.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1)
        if host is not None:
            sr, cr = oracle.ref_ik_batch(rot, trans, free)
            s, c, st = host(rot, trans, free)
            mism = np.nonzero(c != cr)[0]
            unres = int(((st & 2) != 0).sum())
            total += len(c)
            bad += len(mism)
            print(f"{name:28s} solves {len(c):7d} mismatches {len(mism):6d} unresolved-flag {unres:6d} "
                  f"degenerate-flag {int(((st & 1) != 0).sum()):6d}")
            for i in mism[:3]:
                p, f = divmod(int(i), nf)
                print("    q=", np.array2string(q[:, p], precision=17, separator=","), "free=", repr(free[f, p]),
                      "ref", cr[i], "got", c[i], "status", st[i])
    rot, trans, free = wrist_axis_family(a.n_per, a.seed + 1)
    if cov is not None:
        c = np.zeros(rot.shape[1], np.int32)
        cov.ref_ik_batch(rot.shape[1], rot.ctypes.data_as(dp), trans.ctypes.data_as(dp), free.ctypes.data_as(dp),
                         1, 0, None, c.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1)
    if host is not None:
        sr, cr = oracle.ref_ik_batch(rot, trans, free)
        s, c, st = host(rot, trans, free)
        total += len(c)
        bad += int((c != cr).sum())
        print(f"{'wrist_axis':28s} solves {len(c):7d} mismatches {int((c != cr).sum()):6d} unresolved-flag "
              f"{int(((st & 2) != 0).sum()):6d} degenerate-flag {int(((st & 1) != 0).sum()):6d}  "
              f"ref counts {np.bincount(cr, minlength=9).tolist()}")
    if host is not None:
        print(f"TOTAL {total} solves, {bad} mismatches")


if __name__ == "__main__":
    main()
