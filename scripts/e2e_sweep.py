#!/usr/bin/env python
"""e2e (host-buffer) throughput of tcmp_rne_batch_host against the pipeline chunk size (tuning aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import sample_states, N_STATES
from torque_constrained_motion_planning_b200 import engine
pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
hq, hqd, hqdd, hm = (pin(a) for a in sample_states(N_STATES, 2))
htau = torch.empty((7, N_STATES), dtype=torch.float64).pin_memory()
hok = torch.empty((N_STATES,), dtype=torch.uint8).pin_memory()
arrs = [t.numpy() for t in (hq, hqd, hqdd, hm, htau, hok)]
for chunk in [1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20]:
    ws = engine.Workspace(chunk_states=chunk)
    for mode_tau in (True, False):
        f = lambda: engine.torque_test_batch_host_into(ws, "rne", "f64", arrs[0], arrs[1], arrs[2], arrs[3], 0.0, 0.01,
                                                       arrs[4] if mode_tau else None, arrs[5])
        for _ in range(3): f()
        t0 = time.perf_counter()
        K = 10
        for _ in range(K): f()
        dt = (time.perf_counter() - t0) / K
        print("chunk %8d tau=%d  %.3f ms  %.1f M states/s  H2D %.1f GB/s" % (chunk, mode_tau, dt * 1e3, N_STATES / dt / 1e6, 176e6 / dt / 1e9), flush=True)
    ws.close()
# plain cudaMemcpy bandwidth for reference
d = torch.empty_like(hq, device="cuda")
for _ in range(3): d.copy_(hq, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(hq, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("pinned H2D memcpy 56 MB: %.1f GB/s" % (56e6 / dt / 1e9))
h2 = torch.empty_like(hq).pin_memory()
for _ in range(3): h2.copy_(d, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): h2.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("pinned D2H memcpy 56 MB: %.1f GB/s" % (56e6 / dt / 1e9))
