"""Drop-in for the compiled ``ikfast_panda_arm`` CPython extension (ikfast_panda_arm.cpp:12839-12993),
backed by the batched CUDA solver in libtcmp.so.

``get_ik(rot, pos, free)`` / ``get_fk(q)`` keep the extension's list-in / list-out contract, including
``None`` when there is no solution (:12879-12883) and the requirement of exactly three list arguments
(:12854).  ``get_ik_batch`` / ``get_fk_batch`` are the array forms the planner uses.
"""
from __future__ import annotations

import numpy as np

from . import engine


def get_fk(joints):
    """[pos(3), rot(3x3 rows)] of panda_link8 in panda_link0 (ComputeFk, ikfast_panda_arm.cpp:307)."""
    if not isinstance(joints, list):
        raise TypeError("get_fk expects a list")  # PyArg_ParseTuple "O!" with PyList_Type (:12914)
    q = np.asarray(joints, dtype=np.float64).reshape(7, 1)
    trans, rot = engine.fk_batch(q)
    return [trans[:, 0].tolist(), rot[:, 0].reshape(3, 3).tolist()]


def get_ik(rot, pos, free):
    """All IK solutions (list of 7-lists) for rotation ``rot`` (3 rows), position ``pos`` and free
    value ``free=[j7]``, or None -- ComputeIk + solution expansion (:12839-12905)."""
    for a in (rot, pos, free):
        if not isinstance(a, list):
            raise TypeError("get_ik expects three lists")
    r = np.asarray(rot, dtype=np.float64).reshape(9, 1)
    t = np.asarray(pos, dtype=np.float64).reshape(3, 1)
    f = np.asarray(free, dtype=np.float64).reshape(-1)[:1]
    sols, counts, _ = engine.ik_batch(r, t, f)
    c = int(counts[0])
    if c == 0:
        return None
    return sols[0, :min(c, 8)].tolist()


def get_ik_batch(rot9, trans3, free):
    """Array form: rot9 [9][n], trans3 [3][n], free [n_free][n] or [n_free] -> (sols, counts, status)."""
    return engine.ik_batch(rot9, trans3, free)


def get_fk_batch(q):
    return engine.fk_batch(q)
