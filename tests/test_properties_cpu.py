"""CPU: property tests (hypothesis) of the checkers and the host logic -- invariants the domain offers
independently of any fixture: affinity of the torque in the payload mass, equivalence of static calls,
nov == rne at zero velocity, prefix semantics of safe_path with arbitrary predicates, min-jerk boundary values."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from conftest import Q_HI, Q_LO

from torque_constrained_motion_planning_b200 import min_jerk_v2, rrt_star

finite = dict(allow_nan=False, allow_infinity=False)
joint = lambda j: st.floats(float(Q_LO[j]), float(Q_HI[j]), **finite)
conf = st.tuples(*[joint(j) for j in range(7)])
vel = st.tuples(*[st.floats(-2.5, 2.5, **finite) for _ in range(7)])
acc = st.tuples(*[st.floats(-10, 10, **finite) for _ in range(7)])


@settings(max_examples=150, deadline=None)
@given(conf, vel, acc, st.floats(0.02, 8.0, **finite))
def test_torque_is_affine_in_payload_mass(q, qd, qdd, m):
    t0 = oracle.rne(q, qd, qdd, 0.0)
    t1 = oracle.rne(q, qd, qdd, m)
    t2 = oracle.rne(q, qd, qdd, 2 * m)
    assert np.abs((t2 - t0) - 2 * (t1 - t0)).max() < 1e-10


@settings(max_examples=100, deadline=None)
@given(conf, st.floats(0.0, 6.0, **finite))
def test_nov_equals_rne_at_rest_and_ignores_velocities(q, m):
    col = lambda v: np.asarray(v, dtype=float).reshape(7, 1)
    z = np.zeros((7, 1))
    t_nov, ok_nov = oracle.torque_test_batch("nov", col(q), col(np.ones(7)), col(np.ones(7)), m)
    t_rne, ok_rne = oracle.torque_test_batch("rne", col(q), z, z, m)
    t_none, _ = oracle.torque_test_batch("rne", col(q), None, None, m)
    assert np.array_equal(t_nov, t_rne) and np.array_equal(t_none, t_rne) and ok_nov[0] == ok_rne[0]


@settings(max_examples=100, deadline=None)
@given(conf, vel, acc)
def test_joint1_angle_never_matters(q, qd, qdd):
    """Gravity is along joint 1's axis: the torques do not depend on q[0] (the CUDA core relies on this)."""
    q2 = list(q); q2[0] = -q[0] * 0.5 + 0.3
    assert np.abs(oracle.rne(q, qd, qdd, 3.0) - oracle.rne(q2, qd, qdd, 3.0)).max() < 1e-10


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.booleans(), st.booleans()), min_size=0, max_size=40))
def test_safe_path_prefix_semantics(flags):
    """safe_path_force_aware == longest prefix with no collision and torque ok, batched or serial, and the torque
    predicate is never evaluated at or after the first collision (rrt_star.py:92-96)."""
    seq = list(range(len(flags)))
    calls = []
    col = lambda i: flags[i][0]
    def tq(i):
        calls.append(i)
        return flags[i][1]
    expect = []
    for i in seq:
        if flags[i][0] or not flags[i][1]:
            break
        expect.append(i)
    assert rrt_star.safe_path_force_aware(seq, col, tq) == expect
    first_col = next((i for i in seq if flags[i][0]), len(seq))
    assert all(i < first_col for i in calls)
    col_b = lambda i: flags[i][0]
    col_b.batch = lambda s: np.array([flags[i][0] for i in s], dtype=bool)
    seen = []
    tq_b = lambda i: flags[i][1]
    def tq_batch(s):
        seen.extend(s)
        return np.array([flags[i][1] for i in s], dtype=bool)
    tq_b.batch = tq_batch
    assert rrt_star.safe_path_force_aware(seq, col_b, tq_b) == expect
    assert all(i < first_col for i in seen)


@settings(max_examples=60, deadline=None)
@given(st.lists(conf, min_size=2, max_size=6), st.integers(1, 12))
def test_minjerk_hits_every_waypoint_with_zero_end_velocity(points, n):
    pts = np.array(points)
    c = min_jerk_v2.minjerk_coefficients(pts)
    x, v, a = oracle.minjerk_trajectory(min_jerk_v2.coefficients_for_kernel(c), n)
    L = len(points)
    for seg in range(L - 1):
        assert np.abs(x[(seg + 1) * n - 1] - pts[seg + 1]).max() < 1e-9      # t = 1 lands on the next waypoint
    assert np.abs(v[-1]).max() < 1e-9 and np.abs(a[-1]).max() < 1e-8         # last point: gv = ga = 0
