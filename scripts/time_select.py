import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bench import Q_LO, Q_HI
from torque_constrained_motion_planning_b200 import engine
dev = torch.device("cuda")
n, nf = 200_000, 25
rng = np.random.default_rng(3)
q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
free = np.empty((nf, n)); free[0] = q[6]; free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
qd, fd = torch.as_tensor(q, device=dev), torch.as_tensor(free, device=dev)
trans, rot = engine.fk_batch(qd)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print("ik_batch sols   %.2f ms" % t(lambda: engine.ik_batch(rot, trans, fd)))
print("ik_batch counts %.2f ms" % t(lambda: engine.ik_batch(rot, trans, fd, want_sols=False, want_status=False)))
for mode in ("base", "nov", "rne", "dyn"):
    print("ik_select %-4s  %.2f ms" % (mode, t(lambda: engine.ik_select(rot, trans, fd, qd, 3.0, mode=mode))))
fd32 = torch.cat([fd, fd[:7]], 0).contiguous()
print("ik_select rne nf=32 %.2f ms" % t(lambda: engine.ik_select(rot, trans, fd32, qd, 3.0, mode="rne")))
