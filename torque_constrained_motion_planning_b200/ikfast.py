"""Drop-in for the sampling loop of the reference's ``ikfast.py`` (the boundary callers of the IK kernel).

Kept names and call shapes: ``ikfast_inverse_kinematics`` (generator of configurations, ikfast.py:136-169),
``closest_inverse_kinematics`` (:172-188), ``either_inverse_kinematics`` (:208-213), ``is_ik_compiled``,
``import_ikfast``.  The reference reads the current joint values, limits and link frames from PyBullet
(``robot``, ``tool_link``); here ``robot`` is opaque, limits are the Panda constants and the current
configuration is passed as ``current_conf`` (default: the reference's home configuration, utils.py:45).
``world_from_target`` is the pose of ``panda_link8`` in ``panda_link0`` -- the frame pair of PANDA_INFO
(franka_ik_fast.py:19-20) -- as ``(point, quaternion xyzw)``.

The whole free-joint sweep is ONE kernel launch (``ik_utils.ik_sweep``) instead of one ``get_ik`` call per
sampled value.
"""
from __future__ import annotations

import importlib
import random
from itertools import islice

import numpy as np

from .ik_utils import PANDA_INFO, ik_sweep
from .panda_model import Q_LOWER, Q_UPPER, TOP_HOLDING_LEFT_ARM

INF = float("inf")


def get_module_name(ikfast_info=PANDA_INFO):
    return "{}".format(ikfast_info.module_name)   # ikfast.py:24-25


def get_ik_joints(robot=None, ikfast_info=PANDA_INFO, tool_link=None):
    """The 6 + len(free_joints) joints between base and end-effector link (ikfast.py:74-89): the seven arm joints."""
    return list(range(7))


def ikfast_forward_kinematics(robot, ikfast_info, tool_link, conf=None, use_ikfast=True):
    """Pose ``(point, quaternion xyzw)`` of ``panda_link8`` in ``panda_link0`` for ``conf`` (ikfast.py:105-133 with
    world_from_base = tool_from_ee = identity: there is no PyBullet scene to read them from).  ``conf`` defaults
    to the reference's home configuration."""
    from .ik_utils import compute_forward_kinematics
    conf = TOP_HOLDING_LEFT_ARM if conf is None else conf
    return compute_forward_kinematics(import_ikfast(ikfast_info).get_fk, list(conf))


def check_solution(robot, joints, conf, tool_link, target_pose, tolerance=1e-6):
    """ikfast.py:93-102: does FK(conf) reproduce ``target_pose`` (link8 in link0) to ``tolerance``?"""
    from .ik_utils import matrix_from_quat
    point, quat = ikfast_forward_kinematics(robot, PANDA_INFO, tool_link, conf)
    pos_distance = float(np.linalg.norm(np.asarray(point) - np.asarray(target_pose[0])))
    R = np.asarray(matrix_from_quat(quat)).T @ np.asarray(matrix_from_quat(target_pose[1]))
    ori_distance = float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))
    valid = pos_distance <= tolerance and ori_distance <= tolerance
    if not valid:
        print("IKFast warning! | Valid: {} | Position error: {:.3e} | Orientation error: {:.3e}".format(
            valid, pos_distance, ori_distance))
    return valid


def import_ikfast(ikfast_info=PANDA_INFO):
    return importlib.import_module("." + ikfast_info.module_name, package=__package__)


def is_ik_compiled(ikfast_info=PANDA_INFO):
    try:
        from . import _lib
        _lib.load()
        return True
    except Exception as e:  # same contract as the reference: report and return False
        print(e)
        return False


def _length(diff, norm=INF):
    return float(np.linalg.norm(np.asarray(diff, dtype=float), ord=norm))


def ikfast_inverse_kinematics(robot, ikfast_info, tool_link, world_from_target, fixed_joints=[], max_attempts=INF,
                              max_time=INF, norm=INF, max_distance=INF, current_conf=None, rng=None, **kwargs):
    """Yield IK configurations in the order the reference generator does: sweep the free joint (its current
    value first, then uniform samples within its limits, :153-159), shuffle each solve's solutions (:164) and
    keep those inside the joint limits and within ``max_distance`` of the current configuration (:167)."""
    assert (max_attempts < INF) or (max_time < INF)
    if max_distance is None:
        max_distance = INF
    current = np.asarray(TOP_HOLDING_LEFT_ARM if current_conf is None else current_conf, dtype=float)
    attempts = int(max_attempts) if max_attempts < INF else 25
    lower, upper = Q_LOWER.copy(), Q_UPPER.copy()
    if max_distance < INF:   # free joint restricted to current +- max_distance (:150-152)
        lower[6] = max(lower[6], current[6] - max_distance)
        upper[6] = min(upper[6], current[6] + max_distance)
    for conf in ik_sweep(world_from_target, current[6], max_attempts=attempts, rng=rng or random,
                         lower=lower, upper=upper):
        if _length(np.asarray(conf) - current, norm=norm) <= max_distance:
            yield conf


def closest_inverse_kinematics(robot, ikfast_info, tool_link, world_from_target, max_candidates=INF, norm=INF,
                               verbose=True, current_conf=None, **kwargs):
    current = np.asarray(TOP_HOLDING_LEFT_ARM if current_conf is None else current_conf, dtype=float)
    generator = ikfast_inverse_kinematics(robot, ikfast_info, tool_link, world_from_target, norm=norm,
                                          current_conf=current, **kwargs)
    if max_candidates < INF:
        generator = islice(generator, max_candidates)
    solutions = sorted(generator, key=lambda q: _length(np.asarray(q) - current, norm=norm))
    if verbose:
        best = min([INF] + [_length(np.asarray(q) - current, norm=norm) for q in solutions])
        print("Identified {} IK solutions with minimum distance of {:.3f}".format(len(solutions), best))
    return iter(solutions)


def either_inverse_kinematics(robot, ikfast_info, tool_link, world_from_target, fixed_joints=[], use_pybullet=False,
                              **kwargs):
    if use_pybullet:
        raise NotImplementedError("the PyBullet numerical IK fallback (ikfast.py:194-206) is out of scope")
    return closest_inverse_kinematics(robot, ikfast_info, tool_link, world_from_target, fixed_joints=fixed_joints,
                                      **kwargs)
