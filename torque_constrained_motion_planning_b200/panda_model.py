"""Dependency-free Panda model constants (drop-in name for the reference's ``panda_model.py``).

The reference's ``panda_model.Panda`` (panda_model.py:9-57) subclasses roboticstoolbox's ERobot and
loads a URDF from an author-local path; nothing in the reference imports it.  What the hot path
needs from "the robot model" is only numbers, so this module is a plain holder of them: the
modified-DH table (rne.py:47-54), link inertial parameters (rne.py:65-141), joint / velocity /
effort limits (panda_mod.urdf:127-283) and the named configurations the reference class defines
(panda_model.py:49-57).  The same numbers are compiled into libtcmp.so (csrc/panda_model.cuh).
"""
from __future__ import annotations

import math

import numpy as np

PI = math.pi

# (a, d, alpha) per modified-DH row; theta is the joint angle (row 7 is the fixed flange)  rne.py:47-54
DH = np.array([
    [0.0, 0.333, 0.0],
    [0.0, 0.0, -PI / 2],
    [0.0, 0.316, PI / 2],
    [0.0825, 0.0, PI / 2],
    [-0.0825, 0.384, -PI / 2],
    [0.0, 0.0, PI / 2],
    [0.088, 0.0, PI / 2],
    [0.0, 0.107, 0.0],
])

# panda_link1..7, link8, hand  (rne.py:125-136, :106-117, :65-75)
LINK_MASS = np.array([4.970684, 0.646926, 3.228604, 3.587895, 1.225946, 1.666555, 0.735522, 0.0, 0.68])
LINK_COM = np.array([
    [3.875e-03, 2.081e-03, -0.1750], [-3.141e-03, -2.872e-02, 3.495e-03], [2.7518e-02, 3.9252e-02, -6.6502e-02],
    [-5.317e-02, 1.04419e-01, 2.7454e-02], [-1.1953e-02, 4.1065e-02, -3.8437e-02],
    [6.0149e-02, -1.4117e-02, -1.0517e-02], [1.0517e-02, -4.252e-03, 6.1597e-02], [0, 0, 0], [0, 0, 0],
])
# ixx ixy ixz iyy iyz izz about the COM
LINK_INERTIA = np.array([
    [7.0337e-01, -1.3900e-04, 6.7720e-03, 7.0661e-01, 1.9169e-02, 9.1170e-03],
    [7.9620e-03, -3.9250e-03, 1.0254e-02, 2.8110e-02, 7.0400e-04, 2.5995e-02],
    [3.7242e-02, -4.7610e-03, -1.1396e-02, 3.6155e-02, -1.2805e-02, 1.0830e-02],
    [2.5853e-02, 7.7960e-03, -1.3320e-03, 1.9552e-02, 8.6410e-03, 2.8323e-02],
    [3.5549e-02, -2.1170e-03, -4.0370e-03, 2.9474e-02, 2.2900e-04, 8.6270e-03],
    [1.9640e-03, 1.0900e-04, -1.1580e-03, 4.3540e-03, 3.4100e-04, 5.4330e-03],
    [1.2516e-02, -4.2800e-04, -1.1960e-03, 1.0027e-02, -7.4100e-04, 4.8150e-03],
    [0.001, 0.0, 0.0, 0.001, 0.0, 0.001],
    [0.1, 0.0, 0.0, 0.1, 0.0, 0.1],
])

# panda_mod.urdf:127,153,179,205,231,257,283
Q_LOWER = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
Q_UPPER = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
QD_MAX = np.array([2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61])
TAU_MAX = np.array([87.0, 87.0, 87.0, 87.0, 12.0, 12.0, 12.0])

GRAVITY = 9.81             # rne.py:199
PAYLOAD_RADIUS = 0.14 + 0.025  # rne.py:182,186
FLANGE_Z = 0.107           # link7 -> link8
TOOL_Z = 0.105             # hand -> panda_grasptarget (panda_mod.urdf:87-91)
HAND_YAW = -PI / 4         # link8 -> hand (panda_mod.urdf:7-11)

ARM_JOINT_NAMES = ["panda_joint%d" % i for i in range(1, 8)]  # utils.py:29-30
TOP_HOLDING_LEFT_ARM = [0, -PI / 4, 0.0, -6 * PI / 8, 0, PI / 2, PI / 4]  # utils.py:45


class Panda:
    """Constants-only stand-in for ``panda_model.Panda`` (panda_model.py:9): same attribute names for
    the bits the reference class defines (``qdlim``, ``qr``, ``qz``), plus the tables above."""

    name = "panda"
    manufacturer = "Franka Emika"
    n = 7

    def __init__(self):
        self.qdlim = np.array([2.1750, 2.1750, 2.1750, 2.1750, 2.6100, 2.6100, 2.6100, 3.0, 3.0])  # panda_model.py:49-51
        self.qr = np.array([0, -0.3, 0, -2.2, 0, 2.0, np.pi / 4])                                  # :53
        self.qz = np.zeros(7)                                                                      # :54
        self.qlim = np.stack([Q_LOWER, Q_UPPER])
        self.taulim = TAU_MAX.copy()
        self.dh = DH.copy()
        self.configurations = {"qr": self.qr, "qz": self.qz}

    def addconfiguration(self, name, q):
        self.configurations[name] = np.asarray(q, dtype=float)
        setattr(self, name, self.configurations[name])

    def __repr__(self):
        return "Panda(7 DOF, modified DH, limits from panda_mod.urdf)"
