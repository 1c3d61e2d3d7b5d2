"""oracle/rne_numpy_port.py -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the reference's rne.py AT THE REFERENCE'S OWN GRANULARITY: one Python call per state,
4x4 homogeneous DH transforms inverted with np.linalg.inv (rne.py:46-63), 6x6 spatial matrices assembled with
np.block (rne.py:9-27), ten links walked in Python loops (rne.py:217-251).  The C oracle (rne_oracle.c) answers
"what does the reference compute"; this file answers "what does the reference COST per call on this host",
which bench.py reports next to the GPU number because the reference's rne.py itself cannot travel to the GPU box.
Checked against the golden vectors produced by the unmodified rne.py in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

_DH = np.array([[0, 0.333, 0], [0, 0, -np.pi / 2], [0, 0.316, np.pi / 2], [0.0825, 0, np.pi / 2],
                [-0.0825, 0.384, -np.pi / 2], [0, 0, np.pi / 2], [0.088, 0.0, np.pi / 2], [0, 0.107, 0]])
_MASS = [4.970684, 0.646926, 3.228604, 3.587895, 1.225946, 1.666555, 0.735522, 0.0, 0.68]
_COM = np.array([[3.875e-03, 2.081e-03, -0.1750], [-3.141e-03, -2.872e-02, 3.495e-03],
                 [2.7518e-02, 3.9252e-02, -6.6502e-02], [-5.317e-02, 1.04419e-01, 2.7454e-02],
                 [-1.1953e-02, 4.1065e-02, -3.8437e-02], [6.0149e-02, -1.4117e-02, -1.0517e-02],
                 [1.0517e-02, -4.252e-03, 6.1597e-02], [0, 0, 0], [0, 0, 0], [0, 0, 0]])
_INERTIA6 = [[7.0337e-01, -1.3900e-04, 6.7720e-03, 7.0661e-01, 1.9169e-02, 9.1170e-03],
             [7.9620e-03, -3.9250e-03, 1.0254e-02, 2.8110e-02, 7.0400e-04, 2.5995e-02],
             [3.7242e-02, -4.7610e-03, -1.1396e-02, 3.6155e-02, -1.2805e-02, 1.0830e-02],
             [2.5853e-02, 7.7960e-03, -1.3320e-03, 1.9552e-02, 8.6410e-03, 2.8323e-02],
             [3.5549e-02, -2.1170e-03, -4.0370e-03, 2.9474e-02, 2.2900e-04, 8.6270e-03],
             [1.9640e-03, 1.0900e-04, -1.1580e-03, 4.3540e-03, 3.4100e-04, 5.4330e-03],
             [1.2516e-02, -4.2800e-04, -1.1960e-03, 1.0027e-02, -7.4100e-04, 4.8150e-03],
             [0.001, 0.0, 0.0, 0.001, 0.0, 0.001], [0.1, 0.0, 0.0, 0.1, 0.0, 0.1]]


def _sym(i6):
    return np.array([[i6[0], i6[1], i6[2]], [i6[1], i6[3], i6[4]], [i6[2], i6[4], i6[5]]])


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _twist_transform(X):
    R, t = X[:3, :3], X[:3, 3]
    return np.block([[R, _hat(t) @ R], [np.zeros((3, 3)), R]])


def _motion_cross(v):
    w, l = v[3:, 0], v[:3, 0]
    return np.block([[_hat(w), _hat(l)], [np.zeros((3, 3)), _hat(w)]])


def _inertia66(m, c, I):
    C = _hat(c)
    return np.block([[m * np.eye(3), m * C.T], [m * C, I + m * C @ C.T]])


def _child_from_parent(k, theta):
    if k >= 8:
        return np.eye(4)
    a, d, al = _DH[k]
    T = np.array([[np.cos(theta), -np.sin(theta), 0, a],
                  [np.sin(theta) * np.cos(al), np.cos(theta) * np.cos(al), -np.sin(al), -np.sin(al) * d],
                  [np.sin(theta) * np.sin(al), np.cos(theta) * np.sin(al), np.cos(al), np.cos(al) * d],
                  [0, 0, 0, 1]])
    return np.linalg.inv(T)


def rne(q, qd, qdd, payload_mass=0.0):
    """One state, the reference's way.  payload iff payload_mass > 0 (rne.py:184)."""
    nb = 10 if payload_mass > 0 else 9
    q, qd, qdd = (tuple(v) + (0.0, 0.0, 0.0) for v in (q, qd, qdd))
    masses = _MASS + [payload_mass]
    inertias = [_sym(i) for i in _INERTIA6]
    if nb == 10:
        r = 0.14 + 0.025
        inertias.append(np.diag([payload_mass * r * r, payload_mass * r * r, 0.0]))
    vel, acc, wrench, X = [None] * nb, [None] * nb, [None] * nb, [None] * nb
    grav = np.array([[0], [0], [9.81], [0], [0], [0]])
    for k in range(nb):
        vJ = np.array([[0], [0], [0], [0], [0], [qd[k]]])
        aJ = np.array([[0], [0], [0], [0], [0], [qdd[k]]])
        Xk = _child_from_parent(k, q[k])
        if k == 6:
            Xk[2, 3] = 0
        A = _twist_transform(Xk)
        if k == 0:
            vel[k] = vJ
            acc[k] = A @ grav + aJ
        else:
            vel[k] = A @ vel[k - 1] + vJ
            acc[k] = A @ acc[k - 1] + aJ + _motion_cross(vel[k]) @ vJ
        X[k] = Xk
        I = _inertia66(masses[k], _COM[k], inertias[k])
        wrench[k] = I @ acc[k] + (-_motion_cross(vel[k]).T) @ I @ vel[k]
    tau = [0.0] * nb
    for k in range(nb - 1, -1, -1):
        tau[k] = wrench[k][5, 0]
        if k > 0:
            wrench[k - 1] = wrench[k - 1] + _twist_transform(X[k]).T @ wrench[k]
    return np.array(tau[:7])


def torque_test(q, qd, qdd, mass, limits=(87.0, 87.0, 87.0, 87.0, 12.0, 12.0)):
    """panda_primitives.py:171-191 around rne(): payload iff mass > 0.01; infeasible iff any |tau_i| >= limit_i."""
    tau = rne(q, qd, qdd, mass if mass > 0.01 else 0.0)
    return all(abs(tau[i]) < limits[i] for i in range(6)), tau
