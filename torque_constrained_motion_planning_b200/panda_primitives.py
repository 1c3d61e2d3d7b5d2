"""Drop-in for the hot-path half of the reference's ``panda_primitives.py``.

Kept with the reference's names, arguments and return conventions:

* torque-test factories ``get_torque_limits_not_exceded_test_{base, v2, v3_nov, v4}(problem)``
  (panda_primitives.py:13, :60, :118, :155) returning ``test(poses, ptotalMass, velocities,
  accelerations) -> bool``; each closure also has ``.batch(...)`` (many states, one kernel launch),
  ``.mode`` and ``.mass()``.
* ``get_dynamics_fn_v5(problem, resolutions)`` (:295-318) returning ``dynam_fn(path) -> (q, psg, qd, qdd)``,
  plus ``dynam_fn.fused_check(path, torque_fn)`` = min-jerk sampling + torque test of every sample in
  one launch (what rrt_star.py:203-210 does one state at a time).
* ``plan_joint_motion_force_aware`` (:327-346) and ``planner_fn_force_aware(start_conf, pose, problem)``
  (:223-282).  PyBullet is not available (and is out of scope), so ``pose`` is the target pose of the
  ``panda_grasptarget`` frame in the robot base frame, ``problem.fixed`` is a list of
  ``collision.Box`` / ``collision.Sphere`` obstacles, and ``problem.robot`` is opaque.

All torque arithmetic runs in libtcmp.so; nothing here evaluates dynamics on the CPU.
"""
from __future__ import annotations

import datetime
import math
import random

import numpy as np

from . import engine
from . import rne as _rne_mod
from .collision import get_collision_fn
from .ik_utils import ik_sweep, matrix_from_quat, quat_from_matrix
from .min_jerk_v2 import coefficients_for_kernel, minjerk_coefficients, minjerk_trajectory
from .panda_model import HAND_YAW, Q_LOWER, Q_UPPER, TAU_MAX, TOOL_Z
from .rrt_star import rrt_star_force_aware, rrt_star_force_aware_batched
from .utils import (MAX_DISTANCE, SELF_COLLISIONS, check_initial_end_force_aware, create_trajectory,
                    get_arm_joints, get_distance_fn, get_extend_fn, get_mass, get_max_velocities, get_sample_fn)

PI = math.pi
METHOD = "arne"   # panda_primitives.py:8
MASS = 5          # :9


def arm_conf(_, __):
    return [0, -PI / 4, 0.0, -6 * PI / 8, 0, PI / 2, PI / 4]


def _soa(confs):
    a = np.asarray(confs, dtype=np.float64)
    if a.ndim == 1:
        a = a[None, :]
    return np.ascontiguousarray(a[:, :7].T)


def _make_test(problem, mode, mass_fn):
    """Common body of the four factories: scalar closure + batched twin over tcmp_rne_batch.  A ``problem.model``
    (engine.InertialModel; not a reference field) makes every check of the plan use that inertial set."""
    model = getattr(problem, "model", None)

    def batch(confs, ptotalMass=None, velocities=None, accelerations=None):
        q = _soa(confs)
        if mode == "base":
            return np.ones(q.shape[1], dtype=bool)
        m = mass_fn(ptotalMass)
        dyn = mode != "nov" and velocities is not None and accelerations is not None
        qd = _soa(velocities) if dyn else None
        qdd = _soa(accelerations) if dyn else None
        _, ok = engine.torque_test_batch(q, qd, qdd, m, mode=mode, want_tau=False, model=model)
        return ok.astype(bool)

    def test(poses=None, ptotalMass=None, velocities=None, accelerations=None):
        if mode == "base":
            return True
        return bool(batch([poses], ptotalMass, None if velocities is None else [velocities],
                          None if accelerations is None else [accelerations])[0])

    test.batch = batch
    test.mode = mode
    test.mass = lambda: mass_fn(None)
    test.model = model
    test.limits = TAU_MAX.copy() if model is None else model.torque_limit.copy()
    return test


def _problem_mass(problem):
    """totalMass rule of the nov / dyn closures (panda_primitives.py:131-135, :69-74)."""
    m = problem.payload_mass
    if m is None and problem.payload is not None:
        m = get_mass(problem.payload)
    elif problem.payload is None:
        m = 0
    return float(m)


def get_torque_limits_not_exceded_test_base(problem, mass=None):
    return _make_test(problem, "base", lambda p: 0.0)


def get_torque_limits_not_exceded_test_v2(problem, mass=None):
    """`dyn`: tau = M qdd + C qd + g + J^T [0,0,m g,0,0,0] (panda_primitives.py:60-116)."""
    return _make_test(problem, "dyn", lambda p: _problem_mass(problem))


def get_torque_limits_not_exceded_test_v3_nov(problem, mass=None):
    """`nov`: static RNE, velocities and accelerations forced to zero (panda_primitives.py:118-153)."""
    return _make_test(problem, "nov", lambda p: _problem_mass(problem))


def get_torque_limits_not_exceded_test_v4(problem, mass=None):
    """`rne`: full RNE; ptotalMass defaults to problem.payload_mass captured when the closure is built
    (it is a default argument in the reference, panda_primitives.py:171)."""
    captured = problem.payload_mass

    def mass_fn(ptotalMass):
        m = captured if ptotalMass is None else ptotalMass
        if m is None and problem.payload is not None:
            m = get_mass(problem.payload)
        return float(m or 0.0)
    return _make_test(problem, "rne", mass_fn)


_FACTORIES = {"rne": get_torque_limits_not_exceded_test_v4, "dyn": get_torque_limits_not_exceded_test_v2,
              "base": get_torque_limits_not_exceded_test_base, "nov": get_torque_limits_not_exceded_test_v3_nov}


def test_path_torque_constraint(robot, arm, joints, path, mass, r, test_fn):
    """panda_primitives.py:284-293: True iff some configuration of ``path`` EXCEEDS the torque limits.  The reference
    moves the PyBullet robot through the path and calls ``test_fn(arm, ptotalMass=mass, pcomR=r)`` per
    configuration; here the whole path is one batched torque test (``test_fn.batch``) when the test offers it."""
    batch = getattr(test_fn, "batch", None)
    if batch is not None:
        exceeded = not bool(np.asarray(batch(list(path), ptotalMass=mass), dtype=bool).all())
    else:
        exceeded = any(not test_fn(conf, ptotalMass=mass) for conf in path)
    if exceeded:
        print("conf torques exceded in path")
    return exceeded


test_path_torque_constraint.__test__ = False   # a reference-named helper, not a pytest case


def get_dynamics_fn_v5(problem, resolutions):
    num_joints = 7

    def _plan(path):
        m_coeff = minjerk_coefficients(np.array(path))
        move_time = problem.execution_time
        panda_command_freq = 1000  # Hz
        num_intervals = int(move_time * panda_command_freq / len(path))
        return m_coeff, move_time, num_intervals

    def dynam_fn(path, dur=None, vel0=[0.0] * num_joints, acc0=[0.0] * num_joints):
        """Reference return convention (panda_primitives.py:299-316): lists q, psg, qd, qdd."""
        m_coeff, move_time, num_intervals = _plan(path)
        traj = minjerk_trajectory(m_coeff, num_intervals=num_intervals)
        q = [list(p[0]) for p in traj]
        qd = [list(p[1]) for p in traj]
        qdd = [list(p[2]) for p in traj]
        psg = [move_time * n / len(traj) for n in range(0, len(traj))]
        return q, psg, qd, qdd

    def fused_check(path, torque_fn, want_log_torques=False):
        """Samples of the min-jerk trajectory AND their torque test in one kernel launch
        (tcmp_traj_feasibility).  Returns dict(path, vels, accels, psg, feasible, first_fail, tau).  With
        ``dynam_fn.as_arrays = True`` path / vels / accels / psg are NumPy arrays [n][7] instead of the nested
        lists the reference's callers expect (list construction is most of a plan's wall time)."""
        m_coeff, move_time, num_intervals = _plan(path)
        mode = getattr(torque_fn, "mode", "rne")
        mass = torque_fn.mass() if hasattr(torque_fn, "mass") else 0.0
        model = getattr(torque_fn, "model", None)
        out = engine.traj_feasibility(coefficients_for_kernel(m_coeff), num_intervals, mass, mode=mode, model=model)
        n = out["feasible"].shape[0]
        q = out["q"].T.cpu().numpy()
        arrays = getattr(dynam_fn, "as_arrays", False)
        conv = (lambda a: a) if arrays else (lambda a: a.tolist())
        res = {
            "path": conv(q), "vels": conv(out["qd"].T.cpu().numpy()),
            "accels": conv(out["qdd"].T.cpu().numpy()),
            "psg": move_time * np.arange(n) / n if arrays else [move_time * i / n for i in range(n)],
            "feasible": out["first_fail"] == n, "first_fail": out["first_fail"],
            "tau": out["tau"].T.cpu().numpy(),
        }
        if want_log_torques:
            # Conf.__init__'s logging pass (utils.py:3376-3378): rne WITHOUT payload on the same samples
            log = engine.traj_feasibility(coefficients_for_kernel(m_coeff), num_intervals, 0.0, mode="rne",
                                          want_samples=False, model=model)
            res["log_tau"] = log["tau"].T.cpu().numpy()
        return res

    dynam_fn.fused_check = fused_check
    return dynam_fn


def plan_joint_motion_force_aware(body, joints, end_conf, torque_fn, dynam_fn, obstacles=[], attachments=[],
                                  self_collisions=True, disabled_collisions=set(), weights=None, radius=None,
                                  max_distance=MAX_DISTANCE, use_aabb=False, cache=True, custom_limits={},
                                  start_conf=None, collision_fn=None, batch=0, **kwargs):
    """panda_primitives.py:327-346.  ``batch > 0`` switches tree growth to the speculative batched variant
    (rrt_star_force_aware_batched: ``batch`` candidate edges per kernel launch)."""
    assert len(joints) == len(end_conf)
    if (weights is None) and (radius is not None):
        weights = np.reciprocal(radius)
    sample_fn = get_sample_fn(body, joints, custom_limits=custom_limits)
    distance_fn = get_distance_fn(body, joints, weights=weights)
    extend_fn = get_extend_fn(body, joints, resolutions=radius)
    if collision_fn is None:
        collision_fn = get_collision_fn(body, joints, obstacles, attachments, self_collisions, disabled_collisions,
                                        custom_limits=custom_limits)
    if start_conf is None:
        raise ValueError("start_conf is required (there is no simulator to read it from)")
    if not check_initial_end_force_aware(start_conf, end_conf, collision_fn, torque_fn):
        return None, None, None, None
    if batch > 0:
        return rrt_star_force_aware_batched(start_conf, end_conf, weights, sample_fn, radius, collision_fn, torque_fn,
                                            dynam_fn, max_iterations=kwargs.get("max_iterations", 50), batch=batch)
    return rrt_star_force_aware(start_conf, end_conf, distance_fn, sample_fn, extend_fn, collision_fn, torque_fn,
                                dynam_fn, radius=[0.01], **kwargs)


def tool_pose_to_link8(pose):
    """world_from_link8 for a desired world_from_grasptarget: link8 -> hand is Rz(-pi/4)
    (panda_mod.urdf:7-11), hand -> panda_grasptarget is Tz(0.105) (panda_mod.urdf:87-91)."""
    pos, quat = pose
    R = matrix_from_quat(quat)
    c, s = math.cos(HAND_YAW), math.sin(HAND_YAW)
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    R8 = R @ Rz.T
    p8 = np.asarray(pos, dtype=float) - R8 @ (Rz @ np.array([0.0, 0.0, TOOL_Z]))
    return p8, quat_from_matrix(R8)


def bi_panda_inverse_kinematics(robot, arm, gripper_link, gripper_pose, max_attempts=25, max_time=1.3,
                                custom_limits={}, obstacles=[], current_conf=None, collision_fn=None):
    """franka_ik_fast.py:64-79 -> sample_tool_ik (:46-62): first IK solution of the free-joint sweep that is
    inside the joint limits; None on failure or if it collides.  The whole sweep is one kernel launch."""
    current_conf = arm_conf(None, None) if current_conf is None else current_conf
    confs = ik_sweep(tool_pose_to_link8(gripper_pose), current_conf[6], max_attempts=max_attempts)
    for conf in confs[:max_attempts]:
        if np.all(np.asarray(conf) >= Q_LOWER) and np.all(np.asarray(conf) <= Q_UPPER):
            if collision_fn is not None and collision_fn(conf):
                return None
            return tuple(conf)
    return None


def planner_fn_force_aware(start_conf, pose, problem, batch=0, collision_backend="cuda", as_arrays=False):
    """panda_primitives.py:223-282.  Returns a Trajectory (``.path[i].values / .velocities / .accelerations /
    .dt / .torques``) or None.  Extra keyword arguments (defaults keep the reference's behaviour):
    ``batch`` > 0 = speculative batched tree growth; ``collision_backend`` "cuda" | "numpy"; ``as_arrays`` = return
    ``dict(q, qd, qdd, torques, ts)`` of NumPy arrays (the .npz schema of collect_data.py:109-131) instead of
    building one Conf object per sample."""
    timestamp = "{}_{}".format(*str(datetime.datetime.now()).split(" "))
    robot = problem.robot
    obstacles = problem.fixed
    if problem.torque_test not in _FACTORIES:
        raise ValueError("Problem.torque_test must be one of base/dyn/nov/rne, got %r" % (problem.torque_test,))
    torque_test = _FACTORIES[problem.torque_test](problem)
    arm_joints = get_arm_joints(robot)
    resolutions = 0.2 ** np.ones(len(arm_joints))
    dynam_fn = get_dynamics_fn_v5(problem, resolutions)
    dynam_fn.as_arrays = bool(as_arrays)
    grasp = getattr(problem.payload, "grasp", None)
    gripper_pose = pose if grasp is None else grasp(pose)
    collision_fn = get_collision_fn(robot, arm_joints, obstacles, self_collisions=SELF_COLLISIONS,
                                    backend=collision_backend)
    grasp_conf = None
    for _ in range(25):
        grasp_conf = bi_panda_inverse_kinematics(robot, "right", None, gripper_pose, max_attempts=25, max_time=3.5,
                                                 obstacles=obstacles, current_conf=start_conf,
                                                 collision_fn=collision_fn)
        if grasp_conf is not None:
            break
    if grasp_conf is None:
        print("Grasp IK failure", grasp_conf)
        return None
    if not torque_test(grasp_conf):
        print("grasp conf torques exceded")
        return None
    out = plan_joint_motion_force_aware(robot, arm_joints, grasp_conf, torque_test, dynam_fn, obstacles=obstacles,
                                        self_collisions=SELF_COLLISIONS, max_time=50, radius=resolutions / 2,
                                        max_iterations=50, start_conf=start_conf, collision_fn=collision_fn,
                                        batch=batch)
    approach_path, approach_vels, approach_accels, approach_dts = out
    if approach_path is None:
        print("Approach path failure")
        return None
    # torques logged per sample by Conf (utils.py:3376-3378): rne without payload, batched in one launch
    log_tau = _rne_mod.rne_batch(_soa(approach_path), _soa(approach_vels), _soa(approach_accels), 0.0,
                                 model=getattr(problem, "model", None)).T
    if as_arrays:
        return {"q": np.asarray(approach_path), "qd": np.asarray(approach_vels), "qdd": np.asarray(approach_accels),
                "torques": log_tau, "ts": np.asarray(approach_dts)}
    return create_trajectory(robot, arm_joints, approach_path, bodies=[problem.payload], velocities=approach_vels,
                             accelerations=approach_accels, dts=approach_dts, ts=timestamp, torques=log_tau)
