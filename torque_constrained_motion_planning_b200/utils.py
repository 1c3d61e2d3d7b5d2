"""The PyBullet-free subset of the reference's ``utils.py`` that touches the hot path:
``Problem`` (utils.py:86-93), limit getters (:1531-1559), sampling / distance / extend / refine
functions (:2985-3077), ``check_initial_end_force_aware`` (:3323-3338), ``Conf`` / ``Trajectory`` /
``create_trajectory`` (:3340-3414).  The reference versions take a PyBullet body id; here ``body`` is
accepted and ignored (limits are the Panda constants) so call sites keep their shape.
"""
from __future__ import annotations

import math
import random

import numpy as np

from .panda_model import ARM_JOINT_NAMES, Q_LOWER, Q_UPPER, QD_MAX, TAU_MAX, TOP_HOLDING_LEFT_ARM  # noqa: F401

INF = float("inf")
PI = np.pi
MAX_DISTANCE = 0.0
SELF_COLLISIONS = False   # utils.py:56
DEFAULT_RESOLUTION = math.radians(3)


class Problem:
    """utils.py:86-93 verbatim field set.  ``torque_test`` in {"base", "dyn", "nov", "rne"}; the reference
    default "arne" selects nothing and crashes later (panda_primitives.py:242) -- here it raises at once."""

    def __init__(self, robot, fixed, payload, payload_mass, execution_time, torque_test="arne", model=None):
        self.model = model          # not a reference field: engine.InertialModel for a non-stock arm (None = Panda)
        self.robot = robot
        self.fixed = fixed
        self.payload = payload
        self.payload_mass = payload_mass
        self.execution_time = execution_time
        self.torque_test = torque_test


def get_arm_joints(robot=None):
    return list(range(7))


def get_max_force(body, joint):
    return float(TAU_MAX[joint])


def get_max_velocities(body, joints):
    return tuple(float(QD_MAX[j]) for j in joints)


def get_min_limits(body, joints):
    return [float(Q_LOWER[j]) for j in joints]


def get_max_limits(body, joints):
    return [float(Q_UPPER[j]) for j in joints]


def get_mass(payload):
    return float(getattr(payload, "mass", payload if isinstance(payload, (int, float)) else 0.0))


def get_sample_fn(body, joints, custom_limits={}, rng=None, **kwargs):
    """Uniform sampler in the joint limits (utils.py:2985-2990; the reference draws np.random.uniform
    weights per joint through unit_generator)."""
    lower = np.array(get_min_limits(body, joints))
    upper = np.array(get_max_limits(body, joints))
    for j, (lo, hi) in custom_limits.items():
        lower[j], upper[j] = lo, hi
    rng = rng or np.random

    def fn():
        w = rng.uniform(size=len(lower))
        return tuple(w * lower + (1 - w) * upper)   # convex_combination(lower, upper, w)
    return fn


def get_difference_fn(body, joints):
    def fn(q2, q1):
        return tuple(v2 - v1 for v2, v1 in zip(q2, q1))
    return fn


def get_distance_fn(body, joints, weights=None):
    w = np.ones(len(joints)) if weights is None else np.asarray(weights, dtype=float)

    def fn(q1, q2):
        diff = np.asarray(q2, dtype=float) - np.asarray(q1, dtype=float)
        return np.sqrt(np.dot(w, diff * diff))
    return fn


def get_refine_fn(body, joints, num_steps=0):
    """utils.py:3031-3041: num_steps + 1 configurations from (exclusive) q1 to (inclusive) q2."""
    num_steps = num_steps + 1

    def fn(q1, q2):
        q = q1
        for i in range(num_steps):
            positions = (1.0 / (num_steps - i)) * (np.asarray(q2) - np.asarray(q)) + np.asarray(q)
            q = tuple(positions)
            yield q
    return fn


class ExtendSequence:
    """The iterable utils.get_extend_fn returns: iterating it yields exactly the reference's configurations
    (utils.py:3031-3041,3068-3077), and it also remembers its end points and resolutions so that
    rrt_star.safe_path_force_aware can hand the WHOLE edge to the fused kernel (tcmp_extend_prefix) instead of
    materialising it first."""

    def __init__(self, q1, q2, resolutions, norm, body, joints):
        self.q1, self.q2, self.resolutions, self.norm = q1, q2, resolutions, norm
        self._body, self._joints = body, joints

    def _norm(self):
        return np.linalg.norm(np.divide(np.asarray(self.q2) - np.asarray(self.q1), self.resolutions), ord=self.norm)

    def num_configs(self):
        """Number of configurations iteration yields (steps + 1), or None when the norm is not finite."""
        nrm = self._norm()
        return int(nrm) + 1 if np.isfinite(nrm) and nrm < 1.0e6 else None

    def __iter__(self):
        steps = int(self._norm())
        return get_refine_fn(self._body, self._joints, num_steps=steps)(self.q1, self.q2)


def get_extend_fn(body, joints, resolutions=None, norm=2):
    """utils.py:3068-3077: steps = int(|| (q2 - q1) / resolutions ||_norm)."""
    res = DEFAULT_RESOLUTION * np.ones(len(joints)) if resolutions is None else np.asarray(resolutions, dtype=float)

    def fn(q1, q2):
        return ExtendSequence(q1, q2, res, norm, body, joints)
    return fn


def check_initial_end_force_aware(start_conf, end_conf, collision_fn, torque_fn, verbose=True):
    """utils.py:3323-3338."""
    if collision_fn(start_conf):
        print("Warning: initial configuration is in collision")
        return False
    if collision_fn(end_conf):
        print("Warning: end configuration is in collision")
        return False
    if not torque_fn(start_conf):
        print("Warning: initial configuration excedes torque limits")
        return False
    if not torque_fn(end_conf):
        print("Warning: end configuration excedes torque limits")
        return False
    return True


class Conf(object):
    """utils.py:3367-3381.  ``torques`` may be given precomputed (from the fused trajectory kernel) instead
    of being evaluated one state at a time through ``dynam_fn``."""

    def __init__(self, body, joints, values=None, init=False, velocities=None, accelerations=None, movables=None,
                 dt=None, dynam_fn=None, torques=None):
        self.body = body
        self.joints = joints
        self.values = tuple(values)
        self.init = init
        if torques is None and dynam_fn is not None:
            torques = dynam_fn(values, velocities, accelerations)
        self.torques = torques
        self.velocities = velocities[:len(joints)] if velocities is not None else velocities
        self.accelerations = accelerations[:len(joints)] if accelerations is not None else accelerations
        self.dt = dt

    def iterate(self):
        yield self

    def __repr__(self):
        return "q{}".format(id(self) % 1000)


class Trajectory:
    """utils.py:3383-3396 (the forward-direction part; file logging paths are commented out upstream)."""

    def __init__(self, path, bodies=None, ts=None):
        self.path = tuple(path)
        self.bodies = bodies
        self.ts = ts

    def save_npz(self, path):
        """Write the trajectory dump of collect_data.py:109-131 (np.savez with keys q, qd, qdd, torques, ts)."""
        np.savez(path, **self.to_npz_dict())

    def to_npz_dict(self):
        """The on-disk schema of collect_data.py:109-131: q, qd, qdd, torques, ts."""
        return {
            "q": np.array([c.values for c in self.path]),
            "qd": np.array([c.velocities for c in self.path]),
            "qdd": np.array([c.accelerations for c in self.path]),
            "torques": np.array([c.torques for c in self.path]),
            "ts": np.array([c.dt for c in self.path]),
        }


def create_trajectory(robot, joints, path, bodies=None, velocities=None, accelerations=None, movables=None,
                      dts=None, ts=None, dynam_fn=None, torques=None):
    """utils.py:3340-3350.  With ``torques`` ([n][7]) given, no per-sample dynam_fn calls are made."""
    confs = []
    index = 0
    if velocities is not None:
        for i in range(len(velocities)):
            confs.append(Conf(robot, joints, path[i], velocities=velocities[i], movables=bodies,
                              accelerations=accelerations[i], dt=dts[i], dynam_fn=dynam_fn,
                              torques=None if torques is None else torques[i]))
            index += 1
    for i in range(index, len(path)):
        confs.append(Conf(robot, joints, path[i], velocities=None))
    return Trajectory(confs, bodies=bodies, ts=ts)
