"""Structured IK pose families that drive the analytic solver into its singular branches (test data generators).

`structured_families` : joint vectors with joints pinned to special values (0, +-pi/2, +-pi/4, +-3pi/4, +-pi, the joint
limits, joint 4 = +-2.63084142381503) singly, in pairs, mixed, perturbed by 1e-5 .. 1e-9 -- pose = reference
ComputeFk(q); free-joint rows shaped like the reference's sweep (ikfast.py:153-159): the pose's own joint 7 first, then
special and random values.
`wrist_axis_family`   : poses built directly with the shoulder centre on the joint-6 axis (the solver's first guard,
ikfast_panda_arm.cpp:506-508), which no arm configuration reaches.

Used by tests/test_ik_core_host.py (CPU, host build of csrc/ik_core.cuh), tests/test_gpu_fullsize.py (GPU, through the
C ABI) and scripts/ik_structured_families.py (mismatch listing, gcov coverage of the reference).
"""
import math

import numpy as np

Q_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
Q_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
J4_SING = 2.63084142381503     # 2 atan(3.83030303030303): -1 + cos + 3.8303 sin = 0 (ikfast_panda_arm.cpp:512)
SPECIAL = np.array([0.0, math.pi / 2, -math.pi / 2, math.pi / 4, -math.pi / 4, math.pi, -math.pi,
                    3 * math.pi / 4, -3 * math.pi / 4])


def _special_for_joint(j):
    vals = list(SPECIAL) + [Q_LO[j], Q_HI[j]]
    if j == 3:
        vals += [J4_SING, -J4_SING]
    return np.array(vals)


def structured_families(n_per=4000, seed=11, n_free_random=2):
    """dict name -> (q[7][n], free[n_free][n]).  Deterministic for a given (n_per, seed)."""
    rng = np.random.default_rng(seed)
    fams = {}

    def free_rows(q):
        n = q.shape[1]
        rows = [q[6], rng.choice(SPECIAL, size=n), rng.choice(_special_for_joint(6), size=n)]
        for _ in range(n_free_random):
            rows.append(rng.uniform(-3.0, 3.0, size=n))
        return np.stack(rows)

    def rand_q(n):
        return rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))

    # one joint pinned to each of its special values, the others random (in limits)
    for j in range(7):
        vals = _special_for_joint(j)
        q = rand_q(n_per)
        q[j] = vals[np.arange(n_per) % len(vals)]
        fams[f"pin_j{j + 1}"] = (q, free_rows(q))
    # two joints pinned
    for (a, b) in [(0, 1), (1, 2), (1, 3), (3, 4), (3, 5), (4, 5), (5, 6), (1, 5), (2, 4), (1, 4), (0, 2), (3, 6)]:
        q = rand_q(n_per)
        q[a] = rng.choice(_special_for_joint(a), size=n_per)
        q[b] = rng.choice(_special_for_joint(b), size=n_per)
        fams[f"pin_j{a + 1}_j{b + 1}"] = (q, free_rows(q))
    # joint 4 at the elbow-offset singularity, everything else random / special
    for sgn, nm in ((1.0, "p"), (-1.0, "m")):
        q = rand_q(n_per)
        q[3] = sgn * J4_SING
        fams[f"j4_sing_{nm}"] = (q, free_rows(q))
        q = rand_q(n_per)
        q[3] = sgn * J4_SING
        for j in (0, 1, 2, 4, 5, 6):
            m = rng.random(n_per) < 0.4
            q[j, m] = rng.choice(_special_for_joint(j), size=int(m.sum()))
        fams[f"j4_sing_{nm}_special"] = (q, free_rows(q))
    # every joint special with probability 0.3 / 0.6 / 1.0
    for p in (0.3, 0.6, 1.0):
        q = rand_q(n_per)
        for j in range(7):
            m = rng.random(n_per) < p
            q[j, m] = rng.choice(_special_for_joint(j), size=int(m.sum()))
        fams[f"mixed_p{int(p * 10)}"] = (q, free_rows(q))
    # special values outside the joint limits too (the solver does not know the limits)
    q = rng.choice(np.concatenate([SPECIAL, [0.3, -1.2, 2.0, 1.0, -2.5, J4_SING]]), size=(7, n_per))
    fams["grid_unlimited"] = (q, free_rows(q))
    # tiny perturbations of special configurations: straddle the solver's 1e-5 .. 1e-7 thresholds
    for eps in (1e-5, 1e-6, 1e-7, 1e-8):
        q = rand_q(n_per)
        for j in range(7):
            m = rng.random(n_per) < 0.5
            q[j, m] = rng.choice(_special_for_joint(j), size=int(m.sum())) + rng.normal(0, eps, size=int(m.sum()))
        fams[f"near_special_{eps:g}"] = (q, free_rows(q))
    # joint 4 within 1e-5 .. 1e-9 of the elbow singularities (K = 0 at j4 = 0 and j4 = 2.63084142381503): straddles the
    # 1e-6 guard on q0 and the 5e-6 special-angle tests of the solver's fall-back branches
    for centre, nm in ((J4_SING, "sing"), (0.0, "zero"), (-J4_SING, "msing")):
        for eps in (1e-5, 3e-6, 1e-6, 3e-7, 1e-7, 1e-9):
            q = rand_q(n_per // 2)
            q[3] = centre + rng.uniform(-eps, eps, size=q.shape[1])
            for j in (0, 1, 2, 4, 5, 6):
                m = rng.random(q.shape[1]) < 0.25
                q[j, m] = rng.choice(_special_for_joint(j), size=int(m.sum()))
            fams[f"j4_{nm}_pm{eps:g}"] = (q, free_rows(q))
    return fams


def wrist_axis_family(n=4000, seed=12):
    """Poses built directly (not through FK): the shoulder centre within `off` of the joint-6 axis for the given free
    value, i.e. (0.088 - cos(j7) npx + sin(j7) npy, npz) ~ 0 -- the solver's first guard (ikfast_panda_arm.cpp:506-508).
    No arm configuration reaches such a pose (0.384 + 0.316 cos j4 - 0.0825 sin j4 >= 0.057 > 0), so the expected
    count is 0 on every branch the generated tree takes from there, including the doubly singular one where
    |p|^2 also puts joint 4 at 2.63084 (t ~ +-0.068).  Returns (rot9[9][n], trans3[3][n], free[1][n])."""
    rng = np.random.default_rng(seed)
    # random rotations from normalised quaternions
    qt = rng.normal(size=(4, n))
    qt /= np.linalg.norm(qt, axis=0)
    w, x, y, z = qt
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])      # [3][3][n]
    j7 = rng.uniform(-2.9, 2.9, size=n)
    j7[: n // 4] = rng.choice(SPECIAL, size=n // 4)
    t = rng.uniform(-0.8, 0.8, size=n)
    t_sing = math.sqrt((0.986881610513004 + 0.686036892455338 * 0.088 - math.sin(J4_SING - 1.10379390314189))
                       / 3.89793688895078 - 0.088 ** 2)
    k = n // 3
    t[:k] = rng.choice([t_sing, -t_sing], size=k) + rng.choice([0, 1e-9, 1e-7, 1e-6], size=k) * rng.normal(size=k)
    off = rng.choice([0.0, 1e-9, 1e-7, 5e-7, 2e-6, 2e-5, 8e-5, 2e-4], size=(2, n)) * rng.normal(size=(2, n))
    c, s = np.cos(j7), np.sin(j7)
    # np = R^T p with  c npx - s npy = 0.088 + off0,  s npx + c npy = t,  npz = off1
    a = 0.088 + off[0]
    npv = np.array([c * a + s * t, -s * a + c * t, off[1]])
    p = np.einsum("ijn,jn->in", R, npv)
    trans = p + 0.107 * R[:, 2, :] + np.array([0.0, 0.0, 0.333])[:, None]
    rot = R.reshape(9, n)
    return np.ascontiguousarray(rot), np.ascontiguousarray(trans), j7[None, :].copy()
