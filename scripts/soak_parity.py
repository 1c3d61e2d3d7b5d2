#!/usr/bin/env python
"""Soak parity run: many seeded batches through the CUDA path and the CPU checkers, element-by-element.
Reports mask mismatches, max |tau - tau_oracle|, the smallest distance of any torque to its limit, IK count
mismatches against the compiled reference, and edge first-failure mismatches.  One JSON line."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import oracle
from bench import sample_states, Q_LO, Q_HI
from torque_constrained_motion_planning_b200 import engine
NT = len(os.sched_getaffinity(0))
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
B = int(os.environ.get("BATCH", 5_000_000)); NB = int(os.environ.get("BATCHES", 20))
lim = np.array([87.0, 87, 87, 87, 12, 12])[:, None]
out = {"states": 0, "mask_mismatches": 0, "max_tau_err": 0.0, "min_margin": 1e9, "modes": {}}
t0 = time.time()
for b in range(NB):
    q, qd, qdd, m = sample_states(B, seed=1000 + b)
    mode = ("rne", "rne", "nov", "dyn")[b % 4]
    tau_o, ok_o = oracle.torque_test_batch(mode, q, qd, qdd, m, nthreads=NT)
    tau, ok = engine.torque_test_batch(dev(q), dev(qd), dev(qdd), dev(m), mode=mode)
    err = float(np.abs(tau.cpu().numpy() - tau_o).max())
    mism = int((ok.cpu().numpy() != ok_o).sum())
    out["states"] += B; out["mask_mismatches"] += mism; out["max_tau_err"] = max(out["max_tau_err"], err)
    out["min_margin"] = min(out["min_margin"], float(np.abs(lim - np.abs(tau_o[:6])).min()))
    out["modes"][mode] = out["modes"].get(mode, 0) + B
# IK counts
rng = np.random.default_rng(77)
n, nf, rounds = 2_000_000, 10, int(os.environ.get("IK_ROUNDS", 5))
out["ik_solves"] = 0; out["ik_count_mismatches"] = 0; out["ik_status_nonzero"] = 0
for r in range(rounds):
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    free = np.vstack([q[6:7], rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))])
    trans, rot = engine.fk_batch(dev(q))
    _, c, st = engine.ik_batch(rot, trans, dev(free), want_sols=False)
    _, cr = oracle.ref_ik_batch(rot.cpu().numpy(), trans.cpu().numpy(), free, want_sols=False, nthreads=NT)
    out["ik_solves"] += n * nf; out["ik_count_mismatches"] += int((c.cpu().numpy() != cr).sum())
    out["ik_status_nonzero"] += int((st != 0).sum().item())
    stn = st.cpu().numpy()
    out["ik_unflagged_mismatches"] = out.get("ik_unflagged_mismatches", 0) + int(((c.cpu().numpy() != cr) & ((stn & 10) == 0)).sum())
# IK counts on the structured singular-pose families (tests/ik_families.py), a different seed than the test-suite's
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ik_families import structured_families
out["ik_structured_solves"] = 0; out["ik_structured_mismatches"] = 0; out["ik_structured_unflagged_mismatches"] = 0
out["ik_structured_illconditioned"] = 0; out["ik_structured_unresolved"] = 0
for name, (q, free) in structured_families(n_per=int(os.environ.get("IK_FAMILY_N", 60_000)), seed=2024).items():
    trans, rot = oracle.ref_fk_batch(q)
    _, cr = oracle.ref_ik_batch(rot, trans, free, want_sols=False, nthreads=NT)
    _, c, st = engine.ik_batch(dev(rot), dev(trans), dev(free), want_sols=False)
    c, st = c.cpu().numpy(), st.cpu().numpy()
    bad = c != cr
    out["ik_structured_solves"] += len(c); out["ik_structured_mismatches"] += int(bad.sum())
    out["ik_structured_unflagged_mismatches"] += int((bad & ((st & 8) == 0)).sum())
    out["ik_structured_illconditioned"] += int(((st & 8) != 0).sum()); out["ik_structured_unresolved"] += int(((st & 2) != 0).sum())
# edges
E = 1_000_000
qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, E))
qb = np.clip(qa + rng.normal(0, 0.5, size=(7, E)), Q_LO[:, None], Q_HI[:, None])
ff_o = oracle.edge_feasibility("rne", qa, qb, 64, 5.0, nthreads=NT)
ff = engine.edge_feasibility(dev(qa), dev(qb), 64, 5.0)
out["edges"] = E; out["edge_first_fail_mismatches"] = int((ff.cpu().numpy() != ff_o).sum())
out["wall_s"] = time.time() - t0; out["host_threads"] = NT
print(json.dumps(out))
