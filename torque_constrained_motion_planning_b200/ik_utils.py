"""Drop-in for the reference's ``ik_utils.py`` (pose <-> IKFast marshalling), PyBullet-free.

Kept: ``IKFastInfo``, ``USE_ALL`` / ``USE_CURRENT``, ``compute_forward_kinematics(fk_fn, conf)``,
``compute_inverse_kinematics(ik_fn, pose, sampled)`` (returns [] for None, ik_utils.py:29-31) and
``select_solution`` (ik_utils.py:43-52).  Poses are ``(point xyz, quaternion xyzw)`` as in the
reference (utils.py:95-250).  New: ``compute_inverse_kinematics_batch`` and ``ik_sweep`` run whole
free-joint sweeps (ikfast.py:153-169) in one kernel launch.
"""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np

from . import engine
from .panda_model import Q_LOWER, Q_UPPER

IKFastInfo = namedtuple("IKFastInfo", ["module_name", "base_link", "ee_link", "free_joints"])  # ik_utils.py:10

USE_ALL = False
USE_CURRENT = None

PANDA_INFO = IKFastInfo(module_name="ikfast_panda_arm", base_link="panda_link0", ee_link="panda_link8",
                        free_joints=["panda_joint7"])  # franka_ik_fast.py:19-20


def matrix_from_quat(quat):
    """3x3 rotation of a unit quaternion (x, y, z, w) -- what utils.matrix_from_quat gets from
    pybullet.getMatrixFromQuaternion (utils.py:178-179)."""
    x, y, z, w = (float(v) for v in quat)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def quat_from_matrix(R):
    """(x, y, z, w) of a rotation matrix (tf.quaternion_from_matrix semantics, tf.py:1099, re-ordered)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = 2.0 * np.sqrt(1.0 + t)
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = 2.0 * np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def compute_forward_kinematics(fk_fn, conf):
    pos, rot = fk_fn(list(conf))
    return pos, quat_from_matrix(np.array(rot))


def compute_inverse_kinematics(ik_fn, pose, sampled=[]):
    pos, quat = pose
    rot = matrix_from_quat(quat).tolist()
    if len(sampled) == 0:
        solutions = ik_fn(list(rot), list(pos))  # raises TypeError like the extension
    else:
        solutions = ik_fn(list(rot), list(pos), list(sampled))
    if solutions is None:
        return []
    return solutions


def compute_inverse_kinematics_batch(poses, free):
    """poses: sequence of (point, quat); free: [n_free] broadcast or [n_free][n].
    Returns (sols [n*n_free][8][7], counts [n*n_free], status) as NumPy arrays."""
    n = len(poses)
    rot9 = np.empty((9, n))
    trans3 = np.empty((3, n))
    for i, (pos, quat) in enumerate(poses):
        rot9[:, i] = matrix_from_quat(quat).reshape(9)
        trans3[:, i] = pos
    return engine.ik_batch(rot9, trans3, np.asarray(free, dtype=np.float64))


def violates_limits(conf, lower=Q_LOWER, upper=Q_UPPER) -> bool:
    c = np.asarray(conf)
    return bool(np.any(c < lower) or np.any(c > upper))


def ik_sweep(pose, current_free, max_attempts=25, rng=None, lower=Q_LOWER, upper=Q_UPPER):
    """The free-joint sweep of ikfast_inverse_kinematics (ikfast.py:136-169) as ONE launch: free values =
    the current joint-7 value, then uniform samples in its limits (:153-159); per free value the solutions
    are shuffled (:164, utils.randomize) and filtered by joint limits (:167).  Returns the list of
    configurations in the order the reference generator would yield them."""
    rng = rng or random
    free = [float(current_free)] + [rng.uniform(lower[6], upper[6]) for _ in range(max_attempts - 1)]
    sols, counts, _ = compute_inverse_kinematics_batch([pose], np.asarray(free))
    out = []
    for f in range(len(free)):
        confs = sols[f, :min(int(counts[f]), 8)].tolist()
        rng.shuffle(confs)
        out.extend(c for c in confs if not violates_limits(c, lower, upper))
    return out


def get_ik_limits(robot, joint, limits=USE_ALL, current_conf=None):
    """ik_utils.py:34-40: the sampling interval of a free joint -- its URDF limits (USE_ALL), its current value
    (USE_CURRENT; from ``current_conf``, default the home configuration) or the pair given."""
    if limits is USE_ALL:
        return float(Q_LOWER[joint]), float(Q_UPPER[joint])
    if limits is USE_CURRENT:
        from .panda_model import TOP_HOLDING_LEFT_ARM
        value = float((TOP_HOLDING_LEFT_ARM if current_conf is None else current_conf)[joint])
        return value, value
    return limits


def select_solution(body, joints, solutions, nearby_conf=USE_ALL, **kwargs):
    if not solutions:
        return None
    if nearby_conf is USE_ALL:
        return random.choice(solutions)
    if nearby_conf is USE_CURRENT:
        raise ValueError("USE_CURRENT needs a simulator state; pass nearby_conf explicitly")
    ref = np.asarray(nearby_conf)
    return min(solutions, key=lambda conf: float(np.linalg.norm(np.asarray(conf) - ref)))
