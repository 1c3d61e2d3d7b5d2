"""oracle/ref_harness.py -- TEST INFRASTRUCTURE.  Imports the UNMODIFIED Python reference
from /root/reference/src (read-only) so golden vectors can be generated in this container.
It cannot travel to the GPU box: nothing in tests -m gpu, smoke() or bench.py imports this.

Shims (no reference source is edited or copied):
  * ``np.Inf = np.inf`` before ``import rne`` (rne.py:203 uses the alias NumPy 2 removed).
  * ``sys.dont_write_bytecode`` -- the reference directory is read-only.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

REF_SRC = "/root/reference/src"

LIMITS = [87.0, 87.0, 87.0, 87.0, 12.0, 12.0, 12.0]  # panda_mod.urdf:127..283 via utils.py:1558


def available() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "rne.py"))


def _import():
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    import rne as ref_rne  # noqa
    import min_jerk_v2 as ref_mj  # noqa
    return ref_rne, ref_mj


def ref_rne(q, qd, qdd, payload_mass=0.0):
    """rne.rne with add_payload([0,0,0.03], m) first when m > 0 (rne.py:181-198)."""
    R, _ = _import()
    R.remove_payload()
    if payload_mass > 0:
        R.add_payload([0, 0, 0.03], float(payload_mass))
    try:
        return np.asarray(R.rne(list(map(float, q)), list(map(float, qd)), list(map(float, qdd))), dtype=np.float64)
    finally:
        R.remove_payload()


def ref_torque_test(mode, q, qd=None, qdd=None, payload_mass=0.0):
    """The closure bodies of panda_primitives.py:171-191 (rne) / :130-151 (nov) restated around the
    REAL rne.rne -- the closures themselves need utils -> pybullet, which is not installed."""
    R, _ = _import()
    if mode == "base":
        return True, np.zeros(7)
    if mode == "nov" or qd is None or qdd is None:
        qd = [0.0] * 7
        qdd = [0.0] * 7
    R.remove_payload()
    if payload_mass > 0.01:
        R.add_payload([0, 0, 0.03], float(payload_mass))
    try:
        tau = np.asarray(R.rne(list(map(float, q)), list(map(float, qd)), list(map(float, qdd))))
    finally:
        R.remove_payload()
    ok = True
    for i in range(len(LIMITS) - 1):
        if abs(tau[i]) >= LIMITS[i] * 1:
            ok = False
            break
    return ok, tau


class ref_inertial_override:
    """Context manager: overwrite the reference's module-level inertial tables IN PLACE (rne.py:102 inertia_matrices,
    :119 cs, :138 ms are plain lists that rne() re-reads on every call, :212-216) with a flat model record
    (oracle.model_fields layout), and restore them on exit.  No reference code is changed: rne() itself runs as is."""

    def __init__(self, fields):
        self.f = fields

    def __enter__(self):
        R, _ = _import()
        R.remove_payload()
        self.saved = (list(R.ms), [np.array(c, dtype=float) for c in R.cs], [np.array(i) for i in R.inertia_matrices])
        for k in range(9):
            R.ms[k] = float(self.f["mass"][k])
            R.cs[k] = np.array(self.f["com"][k], dtype=float)
            ixx, ixy, ixz, iyy, iyz, izz = (float(v) for v in self.f["inertia"][k])
            R.inertia_matrices[k] = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])  # rne.py:82
        return self

    def __exit__(self, *exc):
        R, _ = _import()
        R.remove_payload()
        R.ms[:], R.cs[:], R.inertia_matrices[:] = self.saved
        return False


def ref_rne_payload_radius(q, qd, qdd, payload_mass, radius):
    """rne.rne after the body of add_payload (rne.py:181-188) run with another lever than hand_width + 0.025:
    the same four calls on the reference's own helpers, with new_inertia([0, 0, radius], m)."""
    R, _ = _import()
    R.remove_payload()
    if payload_mass > 0:
        R.set_has_payload(True)
        R.add_mass_to_ms_global(float(payload_mass))
        R.add_inertia_matrix(R.new_inertia([0, 0, float(radius)], float(payload_mass)))
    try:
        return np.asarray(R.rne(list(map(float, q)), list(map(float, qd)), list(map(float, qdd))), dtype=np.float64)
    finally:
        R.remove_payload()


def ref_minjerk(points, num_intervals):
    """min_jerk_v2.minjerk_coefficients + minjerk_trajectory exactly as get_dynamics_fn_v5 calls them
    (panda_primitives.py:301,310).  Returns (coeffs [k][N][7], x, v, a each [samples][k])."""
    _, M = _import()
    coeffs = M.minjerk_coefficients(np.array(points, dtype=np.float64))
    traj = M.minjerk_trajectory(coeffs, num_intervals=int(num_intervals))
    x = np.array([p[0] for p in traj])
    v = np.array([p[1] for p in traj])
    a = np.array([p[2] for p in traj])
    return coeffs, x, v, a


Q_HOME = [0, -math.pi / 4, 0.0, -6 * math.pi / 8, 0, math.pi / 2, math.pi / 4]  # utils.py:45
