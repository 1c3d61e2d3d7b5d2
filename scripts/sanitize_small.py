#!/usr/bin/env python
"""Every kernel once at small, ragged sizes -- meant for `compute-sanitizer --tool memcheck`, which is CLOSED on this pool
(round 2: "runs under it have left GPUs needing a reset"); there it serves as a plain smoke run and the bounds checks
are the guard-band tests of tests/test_gpu_guards.py."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import sample_states, Q_LO, Q_HI
from torque_constrained_motion_planning_b200 import engine, collision, min_jerk_v2
from torque_constrained_motion_planning_b200.distributed import PeerMaskBuffer
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
n = 1003
q, qd, qdd, m = sample_states(n, 1)
for mode in ("rne", "nov", "dyn", "base"):
    engine.torque_test_batch(dev(q), dev(qd), dev(qdd), dev(m), mode=mode)
    engine.torque_test_batch(dev(q), None, None, 2.0, mode=mode, want_tau=False)
engine.torque_test_batch(dev(q.astype(np.float32)), dev(qd.astype(np.float32)), dev(qdd.astype(np.float32)), 1.0, dtype="f32")
engine.torque_test_batch(q, qd, qdd, m, workspace=engine.Workspace(chunk_states=100))
qb = np.clip(q + 0.3, Q_LO[:, None], Q_HI[:, None])
for W in (1, 33, 64):
    engine.edge_feasibility(dev(q), dev(qb), W, 5.0)
engine.edge_feasibility(q, qb, 17, 5.0, workspace=engine.Workspace(chunk_states=64))
pts = np.random.default_rng(0).uniform(Q_LO, Q_HI, size=(5, 7))
engine.traj_feasibility(min_jerk_v2.coefficients_for_kernel(min_jerk_v2.minjerk_coefficients(pts)), 37, 3.0)
trans, rot = engine.fk_batch(dev(q))
free = np.vstack([q[6:7], np.random.default_rng(1).uniform(-2.8, 2.8, size=(2, n))])
engine.ik_batch(rot, trans, dev(free))
engine.ik_batch(rot, trans, dev(free), want_sols=False, want_status=False)
engine.ik_batch(rot.cpu().numpy(), trans.cpu().numpy(), free, workspace=engine.Workspace(chunk_states=50))
engine.ik_select(rot, trans, dev(free), dev(q), 3.0)
scene = collision.cluttered_scene()
engine.collision_batch(dev(q), scene, payload_radius=0.03)
engine.extend_prefix(dev(q), dev(qb), 0.1 * np.ones(7), scene, 5.0)
# round 2: BASE-mode trajectory, elbow-singular IK solves (redo kernel), completion flags, push, async host batches
engine.traj_feasibility(min_jerk_v2.coefficients_for_kernel(min_jerk_v2.minjerk_coefficients(pts)), 37, 3.0, mode="base")
qs = q.copy(); qs[3] = 2.63084142381503; qs[3, ::3] = 0.0
ts, rs = engine.fk_batch(dev(qs))
fs = np.vstack([qs[6:7], np.zeros((1, n))])
engine.ik_batch(rs, ts, dev(fs))
engine.ik_batch(rs, ts, dev(fs), want_sols=False)
engine.ik_select(rs, ts, dev(fs), dev(qs), 3.0)
buf = PeerMaskBuffer(n)
buf.torque_test(dev(q), dev(qd), dev(qdd), dev(m))
buf.barrier()
for _ in range(4):
    buf.torque_test(dev(q), dev(qd), dev(qdd), dev(m), overlap_gather=True)
buf.join()
_, ok = engine.torque_test_batch(dev(q), dev(qd), dev(qdd), dev(m))
buf.before_step(); buf.push(ok); buf.join()
torch.cuda.synchronize()
buf.close()
pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory().numpy()
ws = engine.Workspace(chunk_states=128)
hq, hqd, hqdd, hm = (pin(a) for a in (q, qd, qdd, m))
ht, ho = pin(np.empty_like(q)), pin(np.empty(n, dtype=np.uint8))
for _ in range(3):
    engine.torque_test_batch_host_async(ws, "rne", "f64", hq, hqd, hqdd, hm, 0.0, 0.01, ht, ho)
engine.workspace_sync(ws)
torch.cuda.synchronize()
print("sanitize_small: all kernels ran")
