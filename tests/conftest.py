import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden():
    return load_golden


# Config-2/4 synthetic distributions; joint/velocity limits from panda_mod.urdf:127..283
Q_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
Q_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
V_LIM = np.array([2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61])


def sample_states(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qd = rng.uniform(-V_LIM[:, None], V_LIM[:, None], size=(7, n))
    qdd = rng.uniform(-10.0, 10.0, size=(7, n))
    mass = rng.choice(np.array([0.0, 1.0, 3.0, 5.0]), size=n)
    return q, qd, qdd, mass


def sample_edges(n, seed):
    rng = np.random.default_rng(seed)
    qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, n)), Q_LO[:, None], Q_HI[:, None])
    return qa, qb
