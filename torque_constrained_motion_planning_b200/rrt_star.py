"""Drop-in for the force-aware half of the reference's ``rrt_star.py`` with batched feasibility checks.

Kept: ``OptimalNode`` (rrt_star.py:18-79), ``safe_path_force_aware`` (:90-98) and
``rrt_star_force_aware`` (:151-211) with the same arguments, return tuple ``(path, vels, accels, psg)``
and ``(None, None, None, None)`` failures.  The reference walks every candidate edge one configuration
at a time -- ``collision(q)`` then ``torque(q)``, one 2.6 ms Python RNE per configuration.  Here an edge
is one batch: if the predicates expose ``.batch`` (ours do), all configurations of the edge go through
one vectorised collision call and ONE torque-kernel launch, and the safe prefix is cut at the first
failure of either -- the same result as the serial loop.

``strict_reference=True`` (default) keeps the reference's observable quirks (with ``informed=True``
that includes its endless loop once a straight-line solution exists -- callers pass ``informed=False``): the time
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
    """First element of ``sequence`` with the smallest ``function`` value (ties keep the earliest, as
    ``scores.index(min(scores))`` does in the reference, rrt_star.py:9-14)."""
    best, best_score = None, None
    for item in sequence:
        score = function(item)
        if best_score is None or score < best_score:
            best, best_score = item, score
    return best


class OptimalNode(object):
    """Tree node with cost-to-come bookkeeping; same attributes and methods as the reference's node
    (rrt_star.py:18-79): ``config``, ``parent``, ``children``, ``d`` (edge length), ``path`` (the intermediate
    configurations of the edge), ``cost``, ``solution``, ``creation``, ``last_rewire``."""

    __slots__ = ("config", "parent", "children", "d", "path", "cost", "solution", "creation", "last_rewire")

    def __init__(self, config, parent=None, d=0, path=[], iteration=None):
        self.config, self.parent, self.d, self.path = config, parent, d, path
        self.children = set()
        self.solution = False
        self.creation = self.last_rewire = iteration
        if parent is None:
            self.cost = d
        else:
            self.cost = parent.cost + d
            parent.children.add(self)

    def set_solution(self, solution):
        node = self
        while node is not None and node.solution is not solution:   # stops where the flag already holds
            node.solution = solution
            node = node.parent

    def retrace(self):
        """Root-to-node list of configurations, edge interiors included."""
        chain = []
        node = self
        while node is not None:
            chain.append(node)
            node = node.parent
        out = []
        for node in reversed(chain):
            out.extend(node.path)
            out.append(node.config)
        return out

    def rewire(self, parent, d, path, iteration=None):
        was_solution = self.solution
        if was_solution:
            self.parent.set_solution(False)
        self.parent.children.discard(self)
        self.parent = parent
        parent.children.add(self)
        if was_solution:
            parent.set_solution(True)
        self.d, self.path, self.last_rewire = d, path, iteration
        self.update()

    def update(self):
        stack = [self]
        while stack:                     # cost-to-come of the whole subtree, iteratively
            node = stack.pop()
            node.cost = node.parent.cost + node.d
            stack.extend(node.children)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, self.config)
    __str__ = __repr__


def elapsed_time(start_time):
    """Seconds since ``start_time`` (rrt_star.py:6-7; the reference's body calls ``time.time()`` on the imported
    FUNCTION and raises AttributeError -- which also kills its plain ``rrt_star`` at the first progress print)."""
    return time() - start_time


def _first_failure(flags_bad):
    flags_bad = np.asarray(flags_bad, dtype=bool)
    return int(np.argmax(flags_bad)) if flags_bad.any() else len(flags_bad)


def _fused_prefix(sequence, collision, torque):
    """One launch for the whole edge when everything needed is on hand: an ExtendSequence (end points +
    resolutions, 2-norm), a CUDA-backed collision predicate (its packed scene) and a torque test from
    panda_primitives (mode + payload).  tcmp_extend_prefix regenerates the same configurations on the device,
    checks collision first and the static torque test only for collision-free ones, and returns the length of
    the safe prefix -- identical to the loop below (tests/test_gpu_parity.py).  Returns None when not applicable."""
    scene = getattr(collision, "scene", None)
    if scene is None or "packed" not in scene or not hasattr(torque, "mode") or not hasattr(sequence, "q1"):
        return None
    if getattr(sequence, "norm", None) != 2:
        return None
    from . import engine
    q1 = np.asarray(sequence.q1, dtype=float).reshape(7, 1)
    q2 = np.asarray(sequence.q2, dtype=float).reshape(7, 1)
    n_host = sequence.num_configs()
    if n_host is None:
        return None              # non-finite end point: let the per-configuration path deal with it
    n_dev, pre = engine.extend_prefix(q1, q2, sequence.resolutions, scene["packed"], torque.mass(), mode=torque.mode,
                                      model=getattr(torque, "model", None))
    # The kernel counts the steps with a sequentially rounded sum, the host (like the reference) with
    # np.linalg.norm, i.e. a BLAS dot with its own ordering; when the norm lands on an integer (axis-aligned moves
    # at resolution 0.1) the two can differ by one step, and a prefix computed for another step count must not
    # truncate this sequence: fall back to the unfused path for that edge.
    if int(n_dev[0]) != n_host:
        return None
    keep = int(pre[0])
    out = []
    for q in sequence:
        if len(out) == keep:
            break
        out.append(q)
    return out


def safe_path_force_aware(sequence, collision, torque):
    """Longest prefix of ``sequence`` whose configurations are collision-free AND within torque limits
    (rrt_star.py:90-98).  Batched when both predicates offer ``.batch``; the torque test is never evaluated
    at or beyond the first collision, exactly like the reference's short-circuit."""
    fused = _fused_prefix(sequence, collision, torque)
    if fused is not None:
        return fused
    configs = list(sequence)
    if not configs:
        return configs
    col_batch, tq_batch = getattr(collision, "batch", None), getattr(torque, "batch", None)
    if col_batch is None or tq_batch is None:
        keep = 0
        for q in configs:
            if collision(q) or not torque(q):
                break
            keep += 1
        return configs[:keep]
    keep = _first_failure(col_batch(configs))
    if keep:
        keep = min(keep, _first_failure(~np.asarray(tq_batch(configs[:keep]), dtype=bool)))
    return configs[:keep]


def safe_path(sequence, collision):
    """Longest collision-free prefix of ``sequence`` (rrt_star.py:82-88); one vectorised call when the predicate
    offers ``.batch``."""
    configs = list(sequence)
    col_batch = getattr(collision, "batch", None)
    if col_batch is not None and configs:
        return configs[:_first_failure(col_batch(configs))]
    keep = 0
    for q in configs:
        if collision(q):
            break
        keep += 1
    return configs[:keep]


def rrt_star(start, goal, distance, sample, extend, collision, radius, max_time=INF, max_iterations=INF,
             goal_probability=.4, informed=True):
    """The torque-blind planner (rrt_star.py:99-149): RRT* over collision-free edges, returns the list of
    configurations or None.  Same sampling / goal-bias / informed-rejection / rewiring rules as the reference's
    source; its progress print (which crashes the reference, see ``elapsed_time``) is dropped, the time bound is
    a real one, rejected samples count as iterations, and both rewiring passes run."""
    if collision(start) or collision(goal):
        return None
    tree = [OptimalNode(start)]
    goal_node = None
    began = time()
    it = 0
    while elapsed_time(began) < max_time and it < max_iterations:
        aim_at_goal = goal_node is None and (it == 0 or random() < goal_probability)
        target = goal if aim_at_goal else sample()
        it += 1     # before the informed rejection: the reference counts only accepted samples (:110-115), so once a
        #             straight-line solution exists every sample is rejected and its loop never ends
        if informed and goal_node is not None and distance(start, target) + distance(target, goal) >= goal_node.cost:
            continue
        nearest = argmin(lambda n: distance(n.config, target), tree)
        grown = safe_path(extend(nearest.config, target), collision)
        if not grown:
            continue
        tip = grown[-1]
        fresh = OptimalNode(tip, parent=nearest, d=distance(nearest.config, tip), path=grown[:-1], iteration=it)
        if aim_at_goal and distance(tip, goal) < 1e-6:
            goal_node = fresh
            goal_node.set_solution(True)
        near = [n for n in tree if np.all(distance(n.config, tip) < radius)]
        tree.append(fresh)
        for n in near:                                   # better parent for the new node?
            d = distance(n.config, tip)
            if n.cost + d < fresh.cost:
                edge = safe_path(extend(n.config, tip), collision)
                if edge and distance(tip, edge[-1]) < 1e-6:
                    fresh.rewire(n, d, edge[:-1], iteration=it)
        for n in near:                                   # is the new node a better parent for its neighbours?
            d = distance(tip, n.config)
            if fresh.cost + d < n.cost:
                edge = safe_path(extend(tip, n.config), collision)
                if edge and distance(n.config, edge[-1]) < 1e-6:
                    n.rewire(fresh, d, edge[:-1], iteration=it)
    return None if goal_node is None else goal_node.retrace()


def rrt_star_force_aware(start, goal, distance, sample, extend, collision, torque_fn, dynam_fn, radius,
                         max_time=INF, max_iterations=INF, goal_probability=.2, informed=False,
                         strict_reference=True):
    """rrt_star.py:151-211.  Returns ``(path, vels, accels, psg)`` or four Nones."""
    failure = (None, None, None, None)
    if collision(start) or collision(goal):
        print("start config in collision")
        return failure
    tree = [OptimalNode(start)]
    goal_node = None
    began = time()
    # the reference's bound is `(t0 - time()) < max_time`, which never trips
    within_time = (lambda: (began - time()) < max_time) if strict_reference else (lambda: (time() - began) < max_time)
    it = 0
    while within_time() and it < max_iterations:
        aim_at_goal = goal_node is None and (it == 0 or random() < goal_probability)
        target = goal if aim_at_goal else sample()
        if informed and goal_node is not None and distance(start, target) + distance(target, goal) >= goal_node.cost:
            continue
        it += 1
        nearest = argmin(lambda n: distance(n.config, target), tree)
        grown = safe_path_force_aware(extend(nearest.config, target), collision, torque_fn)
        if not grown:
            continue
        tip = grown[-1]
        fresh = OptimalNode(tip, parent=nearest, d=distance(nearest.config, tip), path=grown[:-1], iteration=it)
        if aim_at_goal and distance(tip, goal) < 1e-2:
            goal_node = fresh
            goal_node.set_solution(True)
        tree.append(fresh)
        # the reference builds this set lazily, i.e. after the append, so `fresh` is in it (rrt_star.py:183-185)
        near = [n for n in tree if np.all(distance(n.config, fresh.config) < radius)]
        for n in near:                                   # better parent for the new node?
            d = distance(n.config, fresh.config)
            if n.cost + d < fresh.cost:
                edge = safe_path_force_aware(extend(n.config, fresh.config), collision, torque_fn)
                if edge and distance(fresh.config, edge[-1]) < 1e-6:
                    fresh.rewire(n, d, edge[:-1], iteration=it)
        if not strict_reference:                         # the reference's second loop runs on an exhausted iterator
            for n in near:
                if n is fresh:
                    continue
                d = distance(fresh.config, n.config)
                if fresh.cost + d < n.cost:
                    edge = safe_path_force_aware(extend(fresh.config, n.config), collision, torque_fn)
                    if edge and distance(n.config, edge[-1]) < 1e-6:
                        n.rewire(fresh, d, edge[:-1], iteration=it)
    if goal_node is None:
        print("failed to find goal")
        return failure
    waypoints = goal_node.retrace()
    # final check of the smoothed trajectory (rrt_star.py:203-210)
    fused = getattr(dynam_fn, "fused_check", None)
    if fused is not None:
        checked = fused(waypoints, torque_fn)            # min-jerk samples + full torque test: one launch
        if not checked["feasible"]:
            return failure
        return checked["path"], checked["vels"], checked["accels"], checked["psg"]
    path, psg, vels, accels = dynam_fn(waypoints, len(waypoints))
    if path is None:
        return failure
    vels, accels = vels[:len(path)], accels[:len(path)]
    tq_batch = getattr(torque_fn, "batch", None)
    if tq_batch is not None:
        feasible = bool(np.asarray(tq_batch(path, velocities=vels, accelerations=accels), dtype=bool).all())
    else:
        feasible = all(torque_fn(path[i], velocities=vels[i], accelerations=accels[i]) for i in range(len(path)))
    return (path, vels, accels, psg) if feasible else failure


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
                                       scene.get("packed") or scene["obstacles"], torque_fn.mass(),
                                       mode=torque_fn.mode, q_lo=scene["q_lo"], q_hi=scene["q_hi"],
                                       payload_radius=scene["payload_radius"], model=getattr(torque_fn, "model", None))
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
