"""Drop-in for the force-aware half of the reference's ``rrt_star.py`` with batched feasibility checks.

Kept: ``OptimalNode`` (rrt_star.py:18-79), ``safe_path_force_aware`` (:90-98) and
``rrt_star_force_aware`` (:151-211) with the same arguments, return tuple ``(path, vels, accels, psg)``
and ``(None, None, None, None)`` failures.  The reference walks every candidate edge one configuration
at a time -- ``collision(q)`` then ``torque(q)``, one 2.6 ms Python RNE per configuration.  Here an edge
is one batch: if the predicates expose ``.batch`` (ours do), all configurations of the edge go through
one vectorised collision call and ONE torque-kernel launch, and the safe prefix is cut at the first
failure of either -- the same result as the serial loop.

``strict_reference=True`` (default) keeps the reference's observable quirks: the time
bound is ineffective (``t0 - time()`` is never positive, :159), the neighbour set is a one-shot
iterator so only the first rewiring loop ever runs (:183-198), ``radius`` is compared as given
(a one-element list, panda_primitives.py:346).  ``strict_reference=False`` runs both rewiring loops.
"""
from __future__ import annotations

from random import random
from time import time

import numpy as np

INF = float("inf")


def argmin(function, sequence):
    values = list(sequence)
    scores = [function(x) for x in values]
    return values[scores.index(min(scores))]


class OptimalNode(object):
    def __init__(self, config, parent=None, d=0, path=[], iteration=None):
        self.config = config
        self.parent = parent
        self.children = set()
        self.d = d
        self.path = path
        if parent is not None:
            self.cost = parent.cost + d
            self.parent.children.add(self)
        else:
            self.cost = d
        self.solution = False
        self.creation = iteration
        self.last_rewire = iteration

    def set_solution(self, solution):
        if self.solution is solution:
            return
        self.solution = solution
        if self.parent is not None:
            self.parent.set_solution(solution)

    def retrace(self):
        if self.parent is None:
            return self.path + [self.config]
        return self.parent.retrace() + self.path + [self.config]

    def rewire(self, parent, d, path, iteration=None):
        if self.solution:
            self.parent.set_solution(False)
        self.parent.children.remove(self)
        self.parent = parent
        self.parent.children.add(self)
        if self.solution:
            self.parent.set_solution(True)
        self.d = d
        self.path = path
        self.update()
        self.last_rewire = iteration

    def update(self):
        self.cost = self.parent.cost + self.d
        for n in self.children:
            n.update()

    def __str__(self):
        return self.__class__.__name__ + "(" + str(self.config) + ")"
    __repr__ = __str__


def safe_path_force_aware(sequence, collision, torque):
    """Longest prefix of ``sequence`` free of collision and within torque limits (rrt_star.py:90-98)."""
    seq = list(sequence)
    if not seq:
        return []
    col_batch = getattr(collision, "batch", None)
    tq_batch = getattr(torque, "batch", None)
    if col_batch is None or tq_batch is None:
        path = []
        for q in seq:
            if collision(q):
                break
            if not torque(q):
                break
            path.append(q)
        return path
    bad = np.asarray(col_batch(seq), dtype=bool)
    # the reference never evaluates torque past the first collision; neither does this
    stop = int(np.argmax(bad)) if bad.any() else len(seq)
    if stop > 0:
        ok = np.asarray(tq_batch(seq[:stop]), dtype=bool)
        if not ok.all():
            stop = int(np.argmin(ok))
    return seq[:stop]


def rrt_star_force_aware(start, goal, distance, sample, extend, collision, torque_fn, dynam_fn, radius,
                         max_time=INF, max_iterations=INF, goal_probability=.2, informed=False,
                         strict_reference=True):
    if collision(start) or collision(goal):
        print("start config in collision")
        return (None, None, None, None)
    nodes = [OptimalNode(start)]
    goal_n = None
    t0 = time()
    it = 0

    def in_time():
        return (t0 - time()) < max_time if strict_reference else (time() - t0) < max_time

    while in_time() and it < max_iterations:
        do_goal = goal_n is None and (it == 0 or random() < goal_probability)
        s = goal if do_goal else sample()
        if informed and goal_n is not None and distance(start, s) + distance(s, goal) >= goal_n.cost:
            continue
        it += 1

        nearest = argmin(lambda n: distance(n.config, s), nodes)
        path = safe_path_force_aware(extend(nearest.config, s), collision, torque_fn)
        if len(path) == 0:
            continue
        new = OptimalNode(path[-1], parent=nearest, d=distance(nearest.config, path[-1]), path=path[:-1],
                          iteration=it)
        if do_goal and distance(new.config, goal) < 1e-2:
            goal_n = new
            goal_n.set_solution(True)

        nodes.append(new)
        # evaluated after the append, as the reference's lazy filter is (rrt_star.py:183-185)
        neighbors = [n for n in nodes if np.all(distance(n.config, new.config) < radius)]
        for n in neighbors:
            d = distance(n.config, new.config)
            if n.cost + d < new.cost:
                path = safe_path_force_aware(extend(n.config, new.config), collision, torque_fn)
                if len(path) != 0 and distance(new.config, path[-1]) < 1e-6:
                    new.rewire(n, d, path[:-1], iteration=it)
        if not strict_reference:  # the reference's second loop iterates an exhausted filter object
            for n in neighbors:
                if n is new:
                    continue
                d = distance(new.config, n.config)
                if new.cost + d < n.cost:
                    path = safe_path_force_aware(extend(new.config, n.config), collision, torque_fn)
                    if len(path) != 0 and distance(n.config, path[-1]) < 1e-6:
                        n.rewire(new, d, path[:-1], iteration=it)
    if goal_n is None:
        print("failed to find goal")
        return None, None, None, None
    rrt_path = goal_n.retrace()
    # final check on the smoothed trajectory (rrt_star.py:203-210)
    fused = getattr(dynam_fn, "fused_check", None)
    if fused is not None:
        # one launch: min-jerk samples + full RNE torque test of every sample
        out = fused(rrt_path, torque_fn)
        if not out["feasible"]:
            return None, None, None, None
        return out["path"], out["vels"], out["accels"], out["psg"]
    path, psg, vels, accels = dynam_fn(rrt_path, len(rrt_path))
    vels = vels[:len(path)]
    accels = accels[:len(path)]
    if path is None:
        return None, None, None, None
    tq_batch = getattr(torque_fn, "batch", None)
    if tq_batch is not None:
        if not np.asarray(tq_batch(path, velocities=vels, accelerations=accels), dtype=bool).all():
            return None, None, None, None
    else:
        for i in range(len(path)):
            if not torque_fn(path[i], velocities=vels[i], accelerations=accels[i]):
                return None, None, None, None
    return path, vels, accels, psg


def _refine_to(q1, q2, n_steps, k):
    """Configuration index k (0-based) of utils.get_refine_fn(num_steps=n_steps - 1)(q1, q2)."""
    q = np.asarray(q1, dtype=float)
    q2 = np.asarray(q2, dtype=float)
    out = []
    for i in range(k + 1):
        q = (1.0 / (n_steps - i)) * (q2 - q) + q
        out.append(tuple(q))
    return out


def rrt_star_force_aware_batched(start, goal, distance_weights, sample, resolutions, collision, torque_fn, dynam_fn,
                                 max_iterations=50, batch=64, goal_probability=.2, goal_tolerance=1e-2, rng=None):
    """Speculative, batched tree growth: each round draws ``batch`` targets, finds each one's
    nearest tree node on the host (vectorised) and checks ALL candidate edges -- extend steps, collision,
    static torque test, safe prefix -- in ONE kernel launch (tcmp_extend_prefix).  Every edge with a non-empty
    safe prefix adds a node.  Same building blocks and acceptance rules as rrt_star_force_aware (nearest by the
    weighted distance, safe prefix, goal reached when within ``goal_tolerance``), but it consumes the random
    stream in a different order, so its trees differ from the reference's; use strict mode for parity runs.
    ``max_iterations`` counts rounds.  Returns (path, vels, accels, psg) or four Nones."""
    from . import engine
    rng = rng or np.random
    scene = getattr(collision, "scene", None)
    if scene is None or not hasattr(torque_fn, "mode"):
        raise ValueError("batched growth needs collision.get_collision_fn(...) and a torque test from panda_primitives")
    if collision(start) or collision(goal):
        print("start config in collision")
        return (None, None, None, None)
    w = np.asarray(distance_weights, dtype=float)
    res = np.asarray(resolutions, dtype=float)
    goal_a = np.asarray(goal, dtype=float)
    confs = [np.asarray(start, dtype=float)]
    parent = [-1]
    edge = [None]              # (q_from, q_target, n_steps, prefix) to regenerate the intermediate configurations
    goal_idx = None
    for _ in range(max_iterations):
        targets = np.empty((batch, 7))
        is_goal = np.zeros(batch, dtype=bool)
        for b in range(batch):
            if b == 0 or rng.random() < goal_probability:
                targets[b], is_goal[b] = goal_a, True
            else:
                targets[b] = sample()
        C = np.asarray(confs)
        d2 = ((targets[:, None, :] - C[None, :, :]) ** 2 * w).sum(axis=2)
        near = d2.argmin(axis=1)
        q1 = C[near]
        ns, pre = engine.extend_prefix(np.ascontiguousarray(q1.T), np.ascontiguousarray(targets.T), res,
                                       scene["obstacles"], torque_fn.mass(), mode=torque_fn.mode,
                                       q_lo=scene["q_lo"], q_hi=scene["q_hi"], payload_radius=scene["payload_radius"])
        for b in range(batch):
            if pre[b] == 0:
                continue
            new = np.asarray(_refine_to(q1[b], targets[b], int(ns[b]), int(pre[b]) - 1)[-1])
            confs.append(new)
            parent.append(int(near[b]))
            edge.append((q1[b], targets[b], int(ns[b]), int(pre[b])))
            if is_goal[b] and np.sqrt((w * (new - goal_a) ** 2).sum()) < goal_tolerance:
                goal_idx = len(confs) - 1
                break
        if goal_idx is not None:
            break
    if goal_idx is None:
        print("failed to find goal")
        return None, None, None, None
    chain = []
    i = goal_idx
    while i > 0:
        chain.append(i)
        i = parent[i]
    path = [tuple(confs[0])]
    for i in reversed(chain):
        q_from, q_to, n_steps, prefix = edge[i]
        path.extend(_refine_to(q_from, q_to, n_steps, prefix - 1))
    out = dynam_fn.fused_check(path, torque_fn)
    if not out["feasible"]:
        return None, None, None, None
    return out["path"], out["vels"], out["accels"], out["psg"]
