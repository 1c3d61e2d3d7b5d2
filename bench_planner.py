#!/usr/bin/env python
"""bench_planner.py -- end-to-end planner_fn_force_aware wall time (BASELINE.json configs[0] / configs[4]):

  * gpu_strict   : the drop-in planner, reference-identical tree growth (one batch per edge), fused final check
  * gpu_batched  : speculative batched tree growth (tcmp_extend_prefix, 32 candidate edges per launch)
  * cpu_serial   : the SAME planner code with the torque predicate replaced by the CPU oracle called ONE STATE AT
                   A TIME and the NumPy collision stand-in -- the reference planner's structure
                   (rrt_star.py:90-98,203-210) with a torque test ~400x faster than the reference's rne.py
  * reference_measured_s : the reference's OWN planner loop (rrt_star_force_aware + rne.rne + min_jerk_v2 + IKFast,
                   NumPy collision twin, the Conf.__init__ logging sweep included), measured once in the CPU container by
                   scripts/reference_planner_cpu.py with the same seed / scene / start / target and committed as
                   profiles/r02/reference_planner_cpu.json (round 1 quoted an estimate: calls x 2.6 ms)

PyBullet and the reference's scene code are out of scope: all arms use the synthetic scene of
collision.py and the same seeds.  One JSON line per scene.
"""
from __future__ import annotations

import json
import math
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_HOME = [0, -math.pi / 4, 0.0, -3 * math.pi / 4, 0, math.pi / 2, math.pi / 4]
GOAL_Q = [0.7, 0.3, 0.2, -1.9, 0.1, 2.2, 1.0]


def main():
    import torch
    import oracle
    from torque_constrained_motion_planning_b200 import collision, ikfast_panda_arm as ik, ik_utils
    from torque_constrained_motion_planning_b200 import panda_primitives as pp, utils, rrt_star

    pos8, rot8 = ik.get_fk(GOAL_Q)
    c, s = math.cos(-math.pi / 4), math.sin(-math.pi / 4)
    Rt = np.array(rot8) @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    pose = (tuple(np.array(pos8) + Rt @ np.array([0, 0, 0.105])), tuple(ik_utils.quat_from_matrix(Rt)))

    for name, scene, mass, T in [("configs[0]: demo scene, rne, 1 kg, T=5 s", collision.hiro_scene(), 1.0, 5),
                                 ("configs[4]: cluttered scene, rne, 5 kg, T=5 s", collision.cluttered_scene(), 5.0, 5)]:
        def problem():
            return utils.Problem(robot=None, fixed=scene, payload="coke", payload_mass=mass, execution_time=T,
                                 torque_test="rne")

        def timed(fn, reps):
            best, out = None, None
            for r in range(reps):
                random.seed(3); np.random.seed(3)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = fn()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            return best, out

        pp.planner_fn_force_aware(tuple(Q_HOME), pose, problem())          # warm-up (context, first launches)
        t_strict, traj = timed(lambda: pp.planner_fn_force_aware(tuple(Q_HOME), pose, problem()), 3)
        t_batched, traj_b = timed(lambda: pp.planner_fn_force_aware(tuple(Q_HOME), pose, problem(), batch=32), 3)
        t_arrays, _ = timed(lambda: pp.planner_fn_force_aware(tuple(Q_HOME), pose, problem(), batch=32,
                                                              as_arrays=True), 3)

        # CPU arm: same planner, serial per-state predicates
        calls = {"torque": 0}

        def cpu_torque(q, ptotalMass=None, velocities=None, accelerations=None):
            calls["torque"] += 1
            col = lambda v: None if v is None else np.asarray(v, dtype=float)[:7].reshape(7, 1)
            _, ok = oracle.torque_test_batch("rne", col(q), col(velocities), col(accelerations), mass, nthreads=1)
            return bool(ok[0])

        np_col = collision.get_collision_fn(obstacles=scene)
        serial_col = lambda q, verbose=False: np_col(q)

        def cpu_plan():
            p = problem()
            dynam_fn = pp.get_dynamics_fn_v5(p, 0.2 * np.ones(7))
            plain_dynam = lambda path, n=None: dynam_fn(path)               # no fused_check attribute
            grasp = pp.bi_panda_inverse_kinematics(None, "right", None, pose, current_conf=tuple(Q_HOME),
                                                   collision_fn=serial_col)
            if grasp is None or not cpu_torque(grasp):
                return None
            out = pp.plan_joint_motion_force_aware(None, list(range(7)), grasp, cpu_torque, plain_dynam,
                                                   radius=0.1 * np.ones(7), max_iterations=50, max_time=50,
                                                   start_conf=tuple(Q_HOME), collision_fn=serial_col)
            if out[0] is None:
                return None
            # the reference's second sweep: Conf.__init__ evaluates rne (no payload) per sample
            for q, v, a in zip(out[0], out[1], out[2]):
                oracle.rne(q, v, a, 0.0)
                calls["torque"] += 1
            return out

        calls["torque"] = 0
        t_cpu, out_cpu = timed(cpu_plan, 1)
        n_calls = calls["torque"]
        same = (traj is not None and out_cpu is not None and
                np.array(out_cpu[0]).shape == (len(traj.path), 7) and
                np.allclose(np.array([c_.values for c_ in traj.path]), np.array(out_cpu[0]), rtol=0, atol=1e-12))
        try:
            ref = {r["scene"]: r for r in json.load(open(os.path.join(ROOT, "profiles", "r02",
                                                                      "reference_planner_cpu.json")))["results"]}.get(name)
        except Exception:
            ref = None
        ref_s = None if ref is None else ref["reference_measured_s"]
        print(json.dumps({
            "scene": name, "samples": None if traj is None else len(traj.path),
            "gpu_strict_s": t_strict, "gpu_batched_s": t_batched, "gpu_batched_arrays_s": t_arrays,
            "gpu_batched_samples": None if traj_b is None else len(traj_b.path),
            "cpu_serial_s": t_cpu, "cpu_torque_calls": n_calls,
            "reference_measured_s": ref_s, "reference_rne_calls": None if ref is None else ref["rne_calls"],
            "gpu_strict_path_equals_cpu_serial_path": bool(same),
            "speedup_vs_cpu_serial": t_cpu / t_strict,
            "speedup_vs_reference_measured": None if ref_s is None else ref_s / t_strict,
        }))


if __name__ == "__main__":
    main()
