#!/usr/bin/env python
"""GPU box: run the structured singular-pose families through tcmp_ik_batch and dump every solve whose solution count
differs from the compiled reference's (pose, free value, both solution sets) to gpurun_out/ik_mismatch.npz, so the
decision that flipped can be examined on the host build (tests/native/ik_host.cpp)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import oracle
    from ik_families import structured_families
    from torque_constrained_motion_planning_b200 import engine
    n_per = int(os.environ.get("N_PER", 26000))
    NT = len(os.sched_getaffinity(0))
    rows = []
    total = 0
    for name, (q, free) in structured_families(n_per=n_per, seed=int(os.environ.get("SEED", 11))).items():
        trans, rot = oracle.ref_fk_batch(q)
        sr, cr = oracle.ref_ik_batch(rot, trans, free, nthreads=NT)
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
        sols, counts, status = engine.ik_batch(dev(rot), dev(trans), dev(free))
        c = counts.cpu().numpy()
        total += len(c)
        nf = free.shape[0]
        for i in np.nonzero(c != cr)[0]:
            p, f = divmod(int(i), nf)
            rows.append(dict(family=name, q=q[:, p], rot=rot[:, p], trans=trans[:, p], free=free[f, p], ref_count=cr[i],
                             gpu_count=c[i], ref_sols=sr[i], gpu_sols=sols[i].cpu().numpy(),
                             status=int(status[i])))
    print("solves", total, "mismatches", len(rows))
    for r in rows[:40]:
        print(r["family"], "ref", r["ref_count"], "gpu", r["gpu_count"], "status", r["status"],
              np.array2string(r["q"], precision=17, separator=","), repr(float(r["free"])))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "ik_mismatch.npz"),
             **{k: np.array([r[k] for r in rows]) for k in (rows[0].keys() if rows else [])})


if __name__ == "__main__":
    main()
