"""Synthetic collision backend (host logic, NOT the accelerated hot path).

The reference's collision predicate is PyBullet's (utils.get_collision_fn, utils.py:3165-3218), which
is not installed here and is out of scope.  Planner-level runs (BASELINE.json configs
1 and 5) therefore use this stand-in for BOTH planners being compared: link spheres from the DH forward
kinematics against axis-aligned boxes / spheres, plus the joint-limit test the reference's collision_fn
starts with (utils.py:3177-3178).  It is vectorised NumPy over a batch of configurations and exposes the
reference's scalar call ``collision_fn(q) -> bool`` as well as ``collision_fn.batch(qs) -> bool[n]``.
"""
from __future__ import annotations

import numpy as np

from .panda_model import DH, Q_LOWER, Q_UPPER


class Box:
    def __init__(self, center, half_extents, name="box"):
        self.center = np.asarray(center, dtype=float)
        self.half = np.asarray(half_extents, dtype=float)
        self.name = name


class Sphere:
    def __init__(self, center, radius, name="sphere"):
        self.center = np.asarray(center, dtype=float)
        self.radius = float(radius)
        self.name = name


def hiro_scene():
    """Obstacle layout of the reference demo (test_planner.py:36-54): two tables under the arm's base
    plane and a wall behind it, as boxes (table tops at z = 0 with the arm mounted at the origin)."""
    return [
        Box([-0.39905, -0.04297, -0.25], [0.30, 0.45, 0.23], "table_wooden"),
        Box([0.4614, -0.0502, -0.25], [0.30, 0.45, 0.23], "table_ikea"),
        Box([-0.7366, 0.0, 0.6], [0.05, 1.0, 0.9], "wall"),
    ]


def cluttered_scene(n_extra=6, seed=5):
    """Config-5 style clutter: the demo scene plus seeded spheres / boxes in the workspace."""
    rng = np.random.default_rng(seed)
    obs = hiro_scene()
    for i in range(n_extra):
        c = np.array([rng.uniform(0.25, 0.7), rng.uniform(-0.5, 0.5), rng.uniform(0.15, 0.8)])
        if i % 2 == 0:
            obs.append(Sphere(c, rng.uniform(0.04, 0.08), "clutter_sphere%d" % i))
        else:
            obs.append(Box(c, rng.uniform(0.03, 0.07, size=3), "clutter_box%d" % i))
    return obs


def link_frames(qs):
    """Origins [n][9][3] and z axes of the DH frames 0..7 plus the grasp target, for qs [n][7]."""
    qs = np.atleast_2d(np.asarray(qs, dtype=float))
    n = qs.shape[0]
    T = np.tile(np.eye(4), (n, 1, 1))
    origins = np.empty((n, 9, 3))
    for k in range(8):
        a, d, al = DH[k]
        th = qs[:, k] if k < 7 else np.zeros(n)
        c, s, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
        D = np.zeros((n, 4, 4))
        D[:, 0, 0], D[:, 0, 1], D[:, 0, 3] = c, -s, a
        D[:, 1, 0], D[:, 1, 1], D[:, 1, 2], D[:, 1, 3] = s * ca, c * ca, -sa, -sa * d
        D[:, 2, 0], D[:, 2, 1], D[:, 2, 2], D[:, 2, 3] = s * sa, c * sa, ca, ca * d
        D[:, 3, 3] = 1
        T = T @ D
        origins[:, k] = T[:, :3, 3]
    origins[:, 8] = T[:, :3, 3] + 0.105 * T[:, :3, 2]
    return origins


# (frame_from, frame_to, samples, radius): spheres strung along the segments between DH frame origins
_SEGMENTS = [(1, 2, 3, 0.07), (2, 3, 2, 0.07), (3, 4, 4, 0.065), (4, 5, 2, 0.06), (5, 6, 2, 0.055),
             (6, 7, 2, 0.05), (7, 8, 3, 0.05)]


def link_spheres(qs):
    o = link_frames(qs)
    cs, rs = [], []
    for a, b, m, r in _SEGMENTS:
        for t in np.linspace(0.0, 1.0, m):
            cs.append((1 - t) * o[:, a] + t * o[:, b])
            rs.append(r)
    return np.stack(cs, axis=1), np.asarray(rs)   # [n][m][3], [m]


def get_collision_fn(body=None, joints=None, obstacles=(), attachments=(), self_collisions=False,
                     disabled_collisions=(), custom_limits={}, payload_radius=0.0, backend="numpy", **kwargs):
    """backend="numpy": the host reference of the stand-in (below).  backend="cuda": the same predicate in
    libtcmp.so (tcmp_collision_batch); ``collision_fn.scene`` then also feeds the fused edge kernel."""
    lower, upper = Q_LOWER.copy(), Q_UPPER.copy()
    for j, (lo, hi) in custom_limits.items():
        lower[j], upper[j] = lo, hi
    obstacles = list(obstacles)
    if backend == "cuda":
        from . import engine
        packed = engine.PackedScene(obstacles, lower, upper, payload_radius)   # marshalled once per scene

        def batch_cuda(qs):
            qs = np.atleast_2d(np.asarray(qs, dtype=float))
            hit = engine.collision_batch(np.ascontiguousarray(qs[:, :7].T), packed)
            return hit.astype(bool)

        def collision_cuda(q, verbose=False):
            return bool(batch_cuda([q])[0])
        collision_cuda.batch = batch_cuda
        collision_cuda.scene = {"obstacles": obstacles, "q_lo": lower, "q_hi": upper, "payload_radius": payload_radius,
                                "packed": packed}
        return collision_cuda

    def batch(qs):
        qs = np.atleast_2d(np.asarray(qs, dtype=float))
        hit = np.any(qs < lower, axis=1) | np.any(qs > upper, axis=1)      # limits_fn (utils.py:3177)
        if not obstacles:
            return hit
        centers, radii = link_spheres(qs)
        if payload_radius > 0:                                             # held object at the grasp target
            centers = np.concatenate([centers, link_frames(qs)[:, 8:9]], axis=1)
            radii = np.concatenate([radii, [payload_radius]])
        for ob in obstacles:
            if isinstance(ob, Sphere):
                d = np.linalg.norm(centers - ob.center, axis=2)
                hit |= np.any(d < radii + ob.radius, axis=1)
            else:
                delta = np.abs(centers - ob.center) - ob.half
                d = np.linalg.norm(np.maximum(delta, 0.0), axis=2)
                hit |= np.any(d < radii, axis=1)
        return hit

    def collision_fn(q, verbose=False):
        return bool(batch([q])[0])

    collision_fn.batch = batch
    collision_fn.scene = {"obstacles": obstacles, "q_lo": lower, "q_hi": upper, "payload_radius": payload_radius}
    return collision_fn
