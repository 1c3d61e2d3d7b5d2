"""Drop-in for the reference's ``min_jerk_v2.py`` API (piecewise-quintic smoother through RRT waypoints).

``minjerk_coefficients`` keeps the reference's output layout ``[k][N][7]`` (a0..a5 + duration,
min_jerk_v2.py:80-142) including its quirks: unit segment durations by default (:102-103), the
via-velocity rule ``0.5 (v0 + v1) if v0 v1 >= 1e-10 else 0`` (:118), zero velocity at the last point
(:109-110) and an acceleration state that is never carried across segments (:101,132-133).
It is vectorised over joints (tiny host work: N x k scalars).

``minjerk_trajectory`` keeps the list-of-[x, v, a] return of the reference (:144-182) for API
compatibility; the planner does NOT use it on the hot path -- sampling is fused into the CUDA
trajectory kernel (engine.traj_feasibility), which consumes ``coefficients_for_kernel(m_coeffs)``.
"""
from __future__ import annotations

import numpy as np


def minjerk_coefficients(points_array, duration_array=None):
    pts = np.asarray(points_array, dtype=np.float64)
    rows, k = pts.shape
    N = rows - 1
    if duration_array is None:
        duration_array = np.ones(N)
    dur = np.asarray(duration_array, dtype=np.float64)
    assert len(dur) == N, "Invalid number of intervals chosen (must be equal to N+1={})".format(N)
    m = np.zeros((k, N, 7))
    x = pts[0].copy()
    v = np.zeros(k)
    a = np.zeros(k)          # stays zero: the reference never updates it
    for i in range(N):
        gx = pts[i + 1]
        t = dur[i]
        if i == N - 1:
            gv = np.zeros(k)
        else:
            v0 = (pts[i + 1] - pts[i]) / t
            v1 = (pts[i + 2] - pts[i + 1]) / dur[i + 1]
            gv = np.where(v0 * v1 >= 1e-10, 0.5 * (v0 + v1), 0.0)
        ga = np.zeros(k)
        A = (gx - (x + v * t + (a / 2.0) * t * t)) / (t * t * t)
        B = (gv - (v + a * t)) / (t * t)
        C = (ga - a) / t
        m[:, i, 0] = x
        m[:, i, 1] = v
        m[:, i, 2] = a / 2.0
        m[:, i, 3] = 10 * A - 4 * B + 0.5 * C
        m[:, i, 4] = (-15 * A + 7 * B - C) / t
        m[:, i, 5] = (6 * A - 3 * B + 0.5 * C) / (t * t)
        m[:, i, 6] = t
        x, v = gx, gv
    return m


def coefficients_for_kernel(m_coeffs):
    """[k][N][7] (reference layout) -> contiguous [N][k][6] (tcmp_traj_feasibility layout).
    Only unit durations are supported by the kernel (what the planner uses, panda_primitives.py:301)."""
    m = np.asarray(m_coeffs)
    if not np.all(m[:, :, 6] == 1.0):
        raise ValueError("the CUDA trajectory kernel assumes unit segment durations")
    return np.ascontiguousarray(np.transpose(m[:, :, :6], (1, 0, 2)))


def _point(m_coeff, t):
    a0, a1, a2, a3, a4, a5, tm = (m_coeff[:, i] for i in range(7))
    t = t * tm
    x = a0 + a1 * t + a2 * t ** 2 + a3 * t ** 3 + a4 * t ** 4 + a5 * t ** 5
    v = a1 + 2 * a2 * t + 3 * a3 * t ** 2 + 4 * a4 * t ** 3 + 5 * a5 * t ** 4
    a = 2 * a2 + 6 * a3 * t + 12 * a4 * t ** 2 + 20 * a5 * t ** 3
    return x, v, a


_minjerk_trajectory_point = _point


def minjerk_trajectory(m_coeffs, num_intervals, duration_array=None):
    assert num_intervals > 0, "Invalid number of intervals chosen (must be greater than 0)"
    m = np.asarray(m_coeffs)
    n_seg = m.shape[1]
    if duration_array is None:
        duration_array = np.ones(n_seg)
    assert len(duration_array) == n_seg
    ts = np.linspace(1.0 / num_intervals, 1, num_intervals)
    curve = []
    for seg in range(n_seg):
        for t in ts:
            x, v, a = _point(m[:, seg, :], t * duration_array[seg])
            curve.append([x, v, a])
    return curve


def minjerk_point(m_coeffs, m_index, t):
    m = np.asarray(m_coeffs)
    if m_index <= 0:
        return m[:, 0, 0]
    if m_index > m.shape[1]:
        return _point(m[:, m.shape[1] - 1, :], 1)
    t = min(max(t, 0.0), 1.0)
    return _point(m[:, m_index - 1, :], t)
