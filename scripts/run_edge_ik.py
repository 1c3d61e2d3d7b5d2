#!/usr/bin/env python
"""Small driver for ncu: a few launches of the edge kernel (config 4) and the IK kernel (config 3 slice)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import Q_LO, Q_HI
from torque_constrained_motion_planning_b200 import engine
dev = torch.device("cuda")
rng = np.random.default_rng(4)
E, W = 100_000, 64
qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, E))
qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, E)), Q_LO[:, None], Q_HI[:, None])
a, b = torch.as_tensor(qa, device=dev), torch.as_tensor(qb, device=dev)
n, nf = 200_000, 25
q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
free = np.empty((nf, n)); free[0] = q[6]; free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
qd, fd = torch.as_tensor(q, device=dev), torch.as_tensor(free, device=dev)
trans, rot = engine.fk_batch(qd)
for _ in range(4):
    ff = engine.edge_feasibility(a, b, W, 5.0, mode="rne")
    sols, counts, st = engine.ik_batch(rot, trans, fd)
torch.cuda.synchronize()
print("edges feasible", float((ff == W).float().mean()), "ik mean count", float(counts.float().mean()))
