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
