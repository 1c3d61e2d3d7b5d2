#!/usr/bin/env python
"""scripts/reference_planner_cpu.py -- the REFERENCE planner loop, measured (VERDICT r01 missing #2).

Runs, in the CPU container where /root/reference exists, for BASELINE.json configs[0] and configs[4]:

  * the reference's own ``rrt_star.rrt_star_force_aware`` (/root/reference/src/rrt_star.py:151-211), imported
    unmodified;
  * the reference's own ``rne.rne`` / ``add_payload`` / ``remove_payload`` (/root/reference/src/rne.py) behind the
    torque closure of panda_primitives.py:155-193, restated line for line (that module imports PyBullet and cannot be
    imported here): payload iff mass > 0.01, static when no velocities are given, ``|tau_i| >= limit_i`` for joints 0..5;
  * the reference's own ``min_jerk_v2.minjerk_coefficients / minjerk_trajectory`` behind ``get_dynamics_fn_v5``
    (panda_primitives.py:295-318), restated;
  * the Python collision twin of the synthetic scene (collision.get_collision_fn, backend "numpy") -- PyBullet is not
    installed, so both planners see this predicate;
  * the goal configuration from the free-joint sweep (ikfast.py:136-169) over the compiled, unmodified reference
    IKFast (oracle/_ref), one get_ik-equivalent call per free value;
  * the second sweep the reference makes when it builds the trajectory: ``Conf.__init__`` calls ``rne`` once per sample
    without payload (utils.py:3376-3378, panda_primitives.py:281).

Seeds, scene, start configuration and target pose are those of bench.py's planner extras / bench_planner.py, so the
GPU planner's RRT* waypoints can be compared with the ones stored here.  Output: one JSON document (profiles/r02/
reference_planner_cpu.json) with the wall time of every phase, the call counts and the waypoints.

    PYTHONDONTWRITEBYTECODE=1 python scripts/reference_planner_cpu.py > profiles/r02/reference_planner_cpu.json
"""
from __future__ import annotations

import importlib
import json
import math
import os
import platform
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF_SRC = "/root/reference/src"

Q_HOME = [0, -math.pi / 4, 0.0, -3 * math.pi / 4, 0, math.pi / 2, math.pi / 4]      # utils.py:45
GOAL_Q = [0.7, 0.3, 0.2, -1.9, 0.1, 2.2, 1.0]
TAU_LIMITS = [87.0, 87.0, 87.0, 87.0, 12.0, 12.0, 12.0]                                # get_max_force, utils.py:1558
SEED = 3


def import_reference():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("needs the reference tree at /root/reference (CPU container only)")
    np.Inf = np.inf          # rne.py uses the alias NumPy 2 removed
    sys.path.insert(0, REF_SRC)
    mods = {}
    for name in ("rne", "rrt_star", "min_jerk_v2"):
        sys.modules.pop(name, None)
        mods[name] = importlib.import_module(name)
        assert mods[name].__file__.startswith(REF_SRC), mods[name].__file__
    sys.path.remove(REF_SRC)
    return mods["rne"], mods["rrt_star"], mods["min_jerk_v2"]


def main():
    import oracle
    from torque_constrained_motion_planning_b200 import collision, ik_utils, utils
    from torque_constrained_motion_planning_b200 import panda_primitives as pp
    ref_rne, ref_rrt, ref_mj = import_reference()
    assert oracle.have_ref(), "oracle/_ref (compiled reference IKFast) missing: make -C oracle"

    # target pose of the grasp-target frame = FK(GOAL_Q) composed with Rz(-pi/4) Tz(0.105), as bench_planner.py
    trans, rot = oracle.ref_fk_batch(np.array([GOAL_Q]).T)
    R8 = rot[:, 0].reshape(3, 3)
    c, s = math.cos(-math.pi / 4), math.sin(-math.pi / 4)
    Rt = R8 @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    pose = (tuple(trans[:, 0] + Rt @ np.array([0, 0, 0.105])), tuple(ik_utils.quat_from_matrix(Rt)))

    results = []
    for name, scene, mass, T in [("configs[0]: demo scene, rne, 1 kg, T=5 s", collision.hiro_scene(), 1.0, 5),
                                 ("configs[4]: cluttered scene, rne, 5 kg, T=5 s", collision.cluttered_scene(), 5.0, 5)]:
        calls = {"torque": 0, "collision": 0, "ik": 0}

        def torque_test(poses=None, ptotalMass=mass, velocities=None, accelerations=None):
            """panda_primitives.py:171-191 around the real rne.rne."""
            calls["torque"] += 1
            totalMass = ptotalMass
            if velocities is None or accelerations is None:
                velocities = [0] * len(poses)
                accelerations = [0] * len(poses)
            if totalMass > 0.01:
                ref_rne.add_payload([0, 0, 0.03], totalMass)
            torques = ref_rne.rne(poses, velocities, accelerations)
            for i in range(len(TAU_LIMITS) - 1):
                if abs(torques[i]) >= TAU_LIMITS[i]:
                    ref_rne.remove_payload()
                    return False
            ref_rne.remove_payload()
            return True

        np_col = collision.get_collision_fn(obstacles=scene)      # backend "numpy": the Python twin

        def collision_fn(q, verbose=False):
            calls["collision"] += 1
            return np_col(q)

        def dynam_fn(path, dur=None):
            """panda_primitives.py:299-316 around the real min_jerk_v2."""
            m_coeff = ref_mj.minjerk_coefficients(np.array(path))
            num_intervals = T * 1000 / len(path)
            traj = ref_mj.minjerk_trajectory(m_coeff, num_intervals=int(num_intervals))
            q = [list(p[0]) for p in traj]
            qd = [list(p[1]) for p in traj]
            qdd = [list(p[2]) for p in traj]
            psg = [T * n / len(traj) for n in range(0, len(traj))]
            return q, psg, qd, qdd

        def goal_ik():
            """franka_ik_fast.py:46-79 / ikfast.py:136-169: sweep joint 7 (current value first, then uniform in its
            limits), per value one get_ik call on the compiled reference, shuffle, joint-limit filter, first survivor."""
            pos8, quat8 = pp.tool_pose_to_link8(pose)
            R = ik_utils.matrix_from_quat(quat8).reshape(9, 1)
            t = np.asarray(pos8, dtype=float).reshape(3, 1)
            lo, hi = ik_utils.Q_LOWER, ik_utils.Q_UPPER
            free = [float(Q_HOME[6])] + [random.uniform(lo[6], hi[6]) for _ in range(24)]
            for f in free:
                calls["ik"] += 1
                sols, counts = oracle.ref_ik_batch(R, t, np.array([[f]]), nthreads=1)
                confs = sols[0, :min(int(counts[0]), 8)].tolist()
                random.shuffle(confs)
                for conf in confs:
                    if not ik_utils.violates_limits(conf, lo, hi):
                        return None if collision_fn(conf) else tuple(conf)
            return None

        joints = list(range(7))
        radius = (0.2 ** np.ones(7)) / 2
        random.seed(SEED)
        np.random.seed(SEED)
        phases = {}
        t0 = time.perf_counter()
        grasp = goal_ik()
        phases["goal_ik_s"] = time.perf_counter() - t0
        assert grasp is not None
        t0 = time.perf_counter()
        ok_grasp = torque_test(grasp)
        assert ok_grasp
        sample_fn = utils.get_sample_fn(None, joints)
        distance_fn = utils.get_distance_fn(None, joints, weights=np.reciprocal(radius))
        extend_fn = utils.get_extend_fn(None, joints, resolutions=radius)
        assert utils.check_initial_end_force_aware(tuple(Q_HOME), grasp, collision_fn, torque_test)
        out = ref_rrt.rrt_star_force_aware(tuple(Q_HOME), grasp, distance_fn, sample_fn, extend_fn, collision_fn,
                                           torque_test, dynam_fn, radius=[0.01], max_time=50, max_iterations=50)
        phases["rrt_star_and_final_check_s"] = time.perf_counter() - t0
        path, vels, accels, dts = out
        assert path is not None, "reference planner found no path"
        calls_plan = dict(calls)
        t0 = time.perf_counter()
        log = [ref_rne.rne(q, v, a) for q, v, a in zip(path, vels, accels)]     # Conf.__init__, payload removed
        phases["conf_torque_logging_s"] = time.perf_counter() - t0
        total = sum(phases.values())
        results.append({
            "scene": name, "seed": SEED, "payload_mass": mass, "execution_time": T, "samples": len(path),
            "reference_measured_s": total, "phases": phases,
            "rne_calls": calls_plan["torque"] + len(log), "torque_test_calls": calls_plan["torque"],
            "collision_calls": calls_plan["collision"], "ik_calls": calls_plan["ik"],
            "ms_per_rne_call": 1e3 * (phases["rrt_star_and_final_check_s"] + phases["conf_torque_logging_s"])
                               / (calls_plan["torque"] + len(log)),
            "grasp_conf": list(grasp),
            "first_sample": list(path[0]), "last_sample": list(path[-1]),
            "samples_sha": __import__("hashlib").sha256(np.round(np.array(path), 9).tobytes()).hexdigest()[:16],
            "sample_q_every_200": np.array(path)[::200].tolist(),
            "max_abs_log_torque": float(np.abs(np.array(log)).max()),
        })
    doc = {
        "what": "reference rrt_star_force_aware + reference rne.rne + reference min_jerk_v2 + reference IKFast + NumPy "
                "collision twin, single process (the reference planner is serial Python)",
        "host": {"cpu": platform.processor() or platform.machine(), "cores_visible": len(os.sched_getaffinity(0)),
                 "python": platform.python_version(), "numpy": np.__version__},
        "script": "scripts/reference_planner_cpu.py", "results": results,
    }
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
