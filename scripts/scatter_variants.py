#!/usr/bin/env python
"""torchrun, N >= 2: per-step device time of the fused gather's variants (tuning aid for the multi-GPU step).

  E  plain kernel (no scatter at all)
  A  scatter kernel only (peer stores, no completion flags)               -- producer-only, as round 1 timed it
  F  A + tcmp_peer_signal + tcmp_peer_wait in the same stream after every step
  G  A + signal / wait on a side stream under the next step's kernel (PeerMaskBuffer overlap_gather=True)
  M  G with the peer stores replaced by NVSwitch multicast stores (tcmp_rne_batch_scatter_mc, multimem.st)
  H  plain kernel, then on the side stream tcmp_peer_push (copy the local mask block to every rank) + signal / wait
Each: 200 steps captured in one CUDA graph, replayed 20x after a barrier, max over ranks.
(The in-kernel completion-flag tail measured in profiles/r02/scatter_signal_variants.log was removed.)
"""
import os
import sys
import json

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import sample_states, N_STATES  # noqa: E402
from torque_constrained_motion_planning_b200 import engine  # noqa: E402
from torque_constrained_motion_planning_b200.distributed import PeerMaskBuffer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
sets = [tuple(torch.as_tensor(a, device=dev) for a in sample_states(N_STATES, 2 + 1000 * rank + s)) for s in range(4)]
out_tau = torch.empty((7, N_STATES), dtype=torch.float64, device=dev)
out_mask = torch.empty((N_STATES,), dtype=torch.uint8, device=dev)
peer = PeerMaskBuffer(N_STATES, multicast=False)      # unicast peer stores (CUDA IPC mappings)
peer_mc = PeerMaskBuffer(N_STATES, multicast=True)    # NVLS: symmetric memory + multicast address, when the box has it


def var_a(i):
    peer.torque_test(*sets[i % 4], mode="rne", out_tau=out_tau)


def var_f(i):
    peer.torque_test(*sets[i % 4], mode="rne", out_tau=out_tau)
    peer.signal()
    peer.wait()


def var_g(i):
    peer.torque_test(*sets[i % 4], mode="rne", out_tau=out_tau, overlap_gather=True)


masks2 = [torch.empty((N_STATES,), dtype=torch.uint8, device=dev) for _ in range(2)]


def var_h(i):
    peer.before_step()
    _, ok = engine.torque_test_batch(*sets[i % 4], mode="rne", out_tau=out_tau, out_mask=masks2[i % 2])
    peer.push(ok)


def var_m(i):
    peer_mc.torque_test(*sets[i % 4], mode="rne", out_tau=out_tau, overlap_gather=True)


def var_e(i):
    engine.torque_test_batch(*sets[i % 4], mode="rne", out_tau=out_tau, out_mask=out_mask)


def capture(fn, k=200):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(2):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    peer.reset()
    peer_mc.reset()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(k):
            fn(i)
        peer.join()
        peer_mc.join()
    peer.reset()
    peer_mc.reset()
    return g


def timeit(g, reps=20, k=200):
    g.replay()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / (reps * k) * 1e3   # us per step


res = {}
for name, fn in (("E_plain", var_e), ("A_scatter", var_a), ("F_signal_wait_in_stream", var_f),
                 ("G_signal_wait_side_stream", var_g), ("H_unfused_push_side_stream", var_h),
                 ("M_multicast_side_stream" if peer_mc.mc_ptr else "M_multicast_unavailable_unicast", var_m),
                 ("E_plain_again", var_e)):
    res[name] = timeit(capture(fn))
if rank == 0:
    print(json.dumps({"world": world, "lib": os.path.basename(os.environ.get("TCMP_LIB", "libtcmp.so")), "us_per_step": res}))
dist.barrier()
dist.destroy_process_group()
