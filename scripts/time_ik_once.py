#!/usr/bin/env python
"""configs[2] slice for ncu: 400k reachable poses x 25 free values through tcmp_ik_batch, counts only then with sets."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import Q_HI, Q_LO  # noqa: E402
from torque_constrained_motion_planning_b200 import engine  # noqa: E402

rng = np.random.default_rng(3)
n, nf = 400_000, 25
q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
free = np.empty((nf, n))
free[0] = q[6]
free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
qd, fd = torch.as_tensor(q, device="cuda"), torch.as_tensor(free, device="cuda")
trans, rot = engine.fk_batch(qd)
for _ in range(3):
    engine.ik_batch(rot, trans, fd, want_sols=False, want_status=False)
engine.ik_batch(rot[:, :100_000].contiguous(), trans[:, :100_000].contiguous(), fd[:, :100_000].contiguous())
torch.cuda.synchronize()
print("ok")
