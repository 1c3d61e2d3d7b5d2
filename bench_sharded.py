#!/usr/bin/env python
"""bench_sharded.py -- BASELINE.json configs[3] (100k candidate edges x 64 min-jerk waypoints, rne, 5 kg) and
configs[2] (IK: poses x free-joint sweep) sharded over the ranks of a torchrun launch, with the NCCL all-gather
of the first-failure indices / IK counts + solution sets that the north star names.  Rank 0 prints one JSON line
per workload and every rank asserts that the gathered result equals the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_sharded.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import Q_HI, Q_LO  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from torque_constrained_motion_planning_b200 import distributed as D, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- edges: every rank holds the full (seeded) edge set, evaluates its block, all-gathers first_fail -----
    E, W = 100_000, 64
    rng = np.random.default_rng(4)
    qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, E))
    qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, E)), Q_LO[:, None], Q_HI[:, None])
    a, b = torch.as_tensor(qa, device=dev), torch.as_tensor(qb, device=dev)
    ms = timed(lambda: D.sharded_edge_feasibility(a, b, W, 5.0, mode="rne"))
    ff, _ = D.sharded_edge_feasibility(a, b, W, 5.0, mode="rne")
    assert torch.equal(ff, engine.edge_feasibility(a, b, W, 5.0, mode="rne")), "sharded != single-GPU first_fail"
    single = engine.edge_feasibility(a, b, W, 5.0, mode="rne")
    # fused form: every rank's edge kernel stores its block's indices into all ranks' gathered buffers (NVLink peers)
    lo, hi = D.shard_bounds(E, rank, world)
    per = max(D.shard_bounds(E, r, world)[1] - D.shard_bounds(E, r, world)[0] for r in range(world))
    buf = D.PeerIndexBuffer(per)
    a_blk, b_blk = a[:, lo:hi].contiguous(), b[:, lo:hi].contiguous()
    # like for like with the NCCL form (which pays for its synchronisation inside the collective): every step publishes
    # its completion and waits for every rank's (tcmp_peer_signal / tcmp_peer_wait on the buffer's side stream), and the
    # step ends when this rank holds every rank's block
    def fused_step():
        buf.edge_feasibility(a_blk, b_blk, W, 5.0, mode="rne", overlap_gather=True)
        buf.join()
    ms_p2p = timed(fused_step)
    fused_step()
    torch.cuda.synchronize()
    got = torch.cat([buf.gathered[r, : D.shard_bounds(E, r, world)[1] - D.shard_bounds(E, r, world)[0]] for r in range(world)])
    assert torch.equal(got, single), "peer-store gather != single-GPU first_fail"
    if rank == 0:
        print(json.dumps({"workload": "configs[3]: 100k edges x 64 waypoints, rne, 5 kg (strong scaling)", "n_gpus": world,
                          "edges_per_s_fused_p2p": E / (ms_p2p * 1e-3), "ms_fused_p2p": ms_p2p,
                          "edges_per_s_nccl": E / (ms * 1e-3), "ms_nccl": ms,
                          "feasible_fraction": float((ff == W).float().mean().item())}))
    buf.close()

    # ---- IK: poses sharded, counts + solution sets all-gathered --------------------------------------------------
    n, nf = 200_000, 25
    rng = np.random.default_rng(3)
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    free = np.empty((nf, n))
    free[0] = q[6]
    free[1:] = rng.uniform(-2.8973, 2.8973, size=(nf - 1, n))
    trans, rot = engine.fk_batch(torch.as_tensor(q, device=dev))
    fd = torch.as_tensor(free, device=dev)
    ms = timed(lambda: D.sharded_ik(rot, trans, fd), reps=5)
    sols, counts, _ = D.sharded_ik(rot, trans, fd)
    s1, c1, _ = engine.ik_batch(rot, trans, fd)
    assert torch.equal(counts, c1) and torch.equal(sols, s1), "sharded != single-GPU IK"
    # what a planner needs from goal IK is ONE configuration per pose: fused select, gather 56 B per pose
    ref = torch.as_tensor(q, device=dev)
    lo, hi = D.shard_bounds(n, rank, world)

    def select_step():
        best, cost, nv = engine.ik_select(rot[:, lo:hi].contiguous(), trans[:, lo:hi].contiguous(),
                                          fd[:, lo:hi].contiguous(), ref[:, lo:hi].contiguous(), 3.0, mode="rne")
        return D.all_gather_ragged(best.T.contiguous(), n)

    ms_sel = timed(select_step, reps=5)
    best_all = select_step()
    b1, _, _ = engine.ik_select(rot, trans, fd, ref, 3.0, mode="rne")
    assert torch.equal(best_all, b1.T.contiguous()), "sharded != single-GPU IK select"
    if rank == 0:
        print(json.dumps({"workload": "configs[2] slice: 200k poses x 25 free values (strong scaling)", "n_gpus": world,
                          "solves_per_s_sets_gathered": n * nf / (ms * 1e-3), "ms_sets_gathered": ms,
                          "solves_per_s_select_gathered": n * nf / (ms_sel * 1e-3), "ms_select_gathered": ms_sel,
                          "note": "sets: NCCL all-gather of 2.2 GB of [8][7] solution sets; select: fused limits + "
                                  "static torque test + nearest, 56 B per pose gathered"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
