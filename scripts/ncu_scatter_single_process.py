#!/usr/bin/env python
"""One process, two GPUs: K1 in its plain and its scatter form (mask bytes stored into a buffer on GPU 0 AND, over
NVLink, into a buffer on GPU 1) so that ncu -- which must not wrap a multi-rank command -- can capture the fused
compute + gather kernel with its NVLink counters next to the plain kernel.

    ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,lts__t_sectors_srcunit_ltcfabric.sum,\
lts__t_sectors_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none -k regex:rne_batch_kernel \
        --csv --log-file gpurun_out/ncu_scatter_n2.csv python scripts/ncu_scatter_single_process.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import N_STATES, sample_states  # noqa: E402
from torque_constrained_motion_planning_b200 import _lib, engine  # noqa: E402

assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
lib = _lib.load()
q, qd, qdd, m = (torch.as_tensor(a, device="cuda:0") for a in sample_states(N_STATES, 2))
tau = torch.empty((7, N_STATES), dtype=torch.float64, device="cuda:0")
mask = torch.empty(N_STATES, dtype=torch.uint8, device="cuda:0")
own = torch.zeros(2 * N_STATES, dtype=torch.uint8, device="cuda:0")
peer = torch.zeros(2 * N_STATES, dtype=torch.uint8, device="cuda:1")
peer.copy_(own)                         # torch enables peer access between the two devices for this copy
torch.cuda.synchronize()
ptrs = (ctypes.c_void_p * 2)(own.data_ptr(), peer.data_ptr())
for rep in range(6):
    engine.torque_test_batch(q, qd, qdd, m, mode="rne", out_tau=tau, out_mask=mask)
    _lib.check(lib.tcmp_rne_batch_scatter(0, 0, N_STATES, q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), m.data_ptr(), 0.0,
                                          0.01, tau.data_ptr(), 2, ptrs, 0, None))
torch.cuda.synchronize()
assert torch.equal(own[:N_STATES], mask) and torch.equal(peer[:N_STATES].to("cuda:0"), mask)
print("ok: plain and scatter masks agree on both devices")
# the edge kernel with IndexDests (tcmp_edge_feasibility_scatter): configs[3], first-failure indices to both devices
from bench import sample_edges  # noqa: E402
E, W = 100_000, 64
qa, qb = (torch.as_tensor(a, device="cuda:0") for a in sample_edges(E, 4))
own_i = torch.full((E,), -1, dtype=torch.int32, device="cuda:0")
peer_i = torch.full((E,), -1, dtype=torch.int32, device="cuda:1")
peer_i.copy_(own_i)
iptrs = (ctypes.c_void_p * 2)(own_i.data_ptr(), peer_i.data_ptr())
for rep in range(4):
    ff = engine.edge_feasibility(qa, qb, W, 5.0, mode="rne")
    _lib.check(lib.tcmp_edge_feasibility_scatter(0, E, W, qa.data_ptr(), qb.data_ptr(), 5.0, 0.01, 0, 2, iptrs, 0, None))
torch.cuda.synchronize()
assert torch.equal(own_i, ff) and torch.equal(peer_i.to("cuda:0"), ff)
print("ok: plain and scatter first-failure indices agree on both devices")
