#!/usr/bin/env python
"""bench.py -- torque-feasibility states/s (Panda 7-DOF RNE) on N B200s, with the roofline of the
dominant kernel, an end-to-end number through the host-buffer C-ABI call, and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 1 M synthetic Panda states (q, qd, qdd, payload mass in
{0,1,3,5} kg) per GPU, fp64, structure-of-arrays [7][n].  A step = one `rne`-mode pass of the
torque test over the batch (tcmp_rne_batch: joint torques [7][n] + feasibility mask [n]); the `nov`
and `dyn` modes are timed the same way and reported under "modes".  Inputs rotate over 4 distinct
1 M-state sets (4 x 176 MB) so no step finds its inputs in the 126 MB L2.  For N > 1 the states shard
across ranks with no data-path collective (weak scaling: 1 M states per GPU); the per-step NCCL
all-gather of the 1 MB feasibility masks IS inside the timed region.

--impl reference times the CPU implementation of the same path on the host cores: the reference's
rne.py is pure Python/NumPy and cannot travel to the GPU box, so this arm runs the C oracle port of
it (oracle/rne_oracle.c, validated <= 2e-13 N.m against rne.py) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_STATES = 1_000_000
N_SETS = 4
FLOPS_PER_STATE = 1654.0       # BASELINE.md section 4 (rne, payload > 0)
BYTES_PER_STATE = 233.0        # 168 q/qd/qdd + 8 mass in; 56 tau + 1 mask out
METRIC = "torque-feasibility states/sec (Panda 7-DOF RNE)"
UNIT = "states/s"

Q_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
Q_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
V_LIM = np.array([2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61])


def sample_states(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qd = rng.uniform(-V_LIM[:, None], V_LIM[:, None], size=(7, n))
    qdd = rng.uniform(-10.0, 10.0, size=(7, n))
    mass = rng.choice(np.array([0.0, 1.0, 3.0, 5.0]), size=n)
    return q, qd, qdd, mass


def workload_config(n_gpus):
    return {
        "workload": "configs[1]: batched RNE sweep, 1M synthetic Panda states (q, qd, qdd, payload) per GPU, "
                    "fp64, mode rne (tau + mask); nov/dyn under 'modes'",
        "states_per_gpu": N_STATES,
        "layout": "SoA [7][n] fp64",
        "l2": "inputs rotate over %d distinct 1M-state sets (%d MB) > 126 MB L2; no explicit flush"
              % (N_SETS, N_SETS * 176),
        "sharding": "states sharded across %d rank(s), no data-path collective; NCCL all-gather of the "
                    "feasibility masks per step inside the timed region when N > 1: fused into the kernel as "
                    "NVLink peer stores (tcmp_rne_batch_scatter, checked against NCCL all_gather), or "
                    "--gather nccl = NCCL all_gather_into_tensor on a side stream" % n_gpus,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def cpu_baseline_run(n_sample, reps, nthreads=0):
    """C oracle port of rne.py + the limit compare, OpenMP over states, on a bounded sample.
    nthreads = 0 -> every core this process may run on (torchrun exports OMP_NUM_THREADS=1; ignore it)."""
    import oracle
    if nthreads == 0:
        nthreads = len(os.sched_getaffinity(0))
    q, qd, qdd, mass = sample_states(n_sample, seed=2)
    oracle.torque_test_batch("rne", q[:, :1000], qd[:, :1000], qdd[:, :1000], mass[:1000], nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.torque_test_batch("rne", q, qd, qdd, mass, nthreads=nthreads)
    dt = time.perf_counter() - t0
    cores = oracle.num_threads() if nthreads == 0 else nthreads
    return n_sample * reps / dt, cores, dt


FP64_INSTR_PER_STATE = 669.0          # 419 DFMA + 180 DMUL + 70 DADD in K1's loop body (SASS)
EXEC_FLOPS_PER_STATE = 2 * 419.0 + 180.0 + 70.0


def sample_edges(n, seed):
    """configs[3] distribution: q_a uniform in the joint limits, q_b = clip(q_a + N(0, 0.5^2))."""
    rng = np.random.default_rng(seed)
    qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, n)), Q_LO[:, None], Q_HI[:, None])
    return qa, qb


def cpu_baseline_extras(nthreads=0):
    """CPU arms of the two other workloads, on bounded samples: the UNMODIFIED reference IKFast (oracle/_ref, compiled
    from the reference's own ikfast_panda_arm.cpp) on 40k poses x 25 free values of configs[2], and the C oracle's
    edge check on 4k edges x 64 waypoints of configs[3]; OpenMP over every host thread."""
    import oracle
    if nthreads == 0:
        nthreads = len(os.sched_getaffinity(0))
    out = {}
    if oracle.have_ref():
        rng = np.random.default_rng(3)
        n, n_free = 40_000, 25
        q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
        free = np.empty((n_free, n))
        free[0] = q[6]
        free[1:] = rng.uniform(-2.8973, 2.8973, size=(n_free - 1, n))
        trans, rot = oracle.ref_fk_batch(q)
        oracle.ref_ik_batch(rot, trans, free, nthreads=nthreads, want_sols=False)      # warm: threads, pages
        t0 = time.perf_counter()
        oracle.ref_ik_batch(rot, trans, free, nthreads=nthreads, want_sols=False)
        dt = time.perf_counter() - t0
        out["ik"] = {"value": n * n_free / dt, "unit": "solves/s", "cores": nthreads, "kind": "reference",
                     "sample": "40k poses x 25 free values (1M solves) of configs[2], ComputeIk of the unmodified "
                               "ikfast_panda_arm.cpp compiled as oracle/_ref, solution counts kept"}
    qa, qb = sample_edges(4_000, 4)
    oracle.edge_feasibility("rne", qa[:, :64], qb[:, :64], 64, 5.0, nthreads=nthreads)
    t0 = time.perf_counter()
    oracle.edge_feasibility("rne", qa, qb, 64, 5.0, nthreads=nthreads)
    dt = time.perf_counter() - t0
    out["edges"] = {"value": 4_000 / dt, "unit": "edges/s", "cores": nthreads, "kind": "port",
                    "sample": "4k edges x 64 min-jerk waypoints of configs[3], rne, 5 kg, C oracle"}
    return out


def python_port_rate(n=300):
    """States/s of ONE core running the NumPy restatement that keeps rne.py's per-call structure (np.block 6x6
    algebra, np.linalg.inv per link): what the reference's own torque test costs on this host."""
    from oracle import rne_numpy_port as P
    q, qd, qdd, mass = sample_states(n, seed=2)
    P.torque_test(q[:, 0], qd[:, 0], qdd[:, 0], mass[0])
    t0 = time.perf_counter()
    for i in range(n):
        P.torque_test(q[:, i], qd[:, i], qdd[:, i], mass[i])
    return n / (time.perf_counter() - t0)


def run_extras(engine, dev, K):
    """IK solves/s (config 3: 1M reachable poses x 25 free values) and RRT* edge checks/s (config 4: 100k edges
    x 64 min-jerk waypoints, rne, 5 kg), device-resident, CUDA-event timed."""
    import torch
    out = {}
    rng = np.random.default_rng(3)
    n, n_free = 1_000_000, 25
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    free = np.empty((n_free, n))
    free[0] = q[6]
    free[1:] = rng.uniform(-2.8973, 2.8973, size=(n_free - 1, n))
    qd_, fd = torch.as_tensor(q, device=dev), torch.as_tensor(free, device=dev)

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    t_fk = timeit(lambda: engine.fk_batch(qd_), 5)
    trans, rot = engine.fk_batch(qd_)
    t_cnt = timeit(lambda: engine.ik_batch(rot, trans, fd, want_sols=False, want_status=False), 3)
    ns = 200_000   # with the [8][7] solution sets written (448 B/solve): 5M solves = 2.2 GB of output
    t_sol = timeit(lambda: engine.ik_batch(rot[:, :ns].contiguous(), trans[:, :ns].contiguous(),
                                           fd[:, :ns].contiguous(), want_status=False), 3)
    _, counts, _ = engine.ik_batch(rot, trans, fd, want_sols=False, want_status=False)
    hist = torch.bincount(counts.flatten().long(), minlength=9).tolist()
    out["ik"] = {"workload": "configs[2]: 1M reachable poses x 25 free-joint values", "solves": n * n_free,
                 "solves_per_s_counts_only": n * n_free / t_cnt, "solves_per_s_with_solutions": ns * n_free / t_sol,
                 "fk_poses_per_s": n / t_fk, "count_histogram": hist}
    E, W = 100_000, 64
    rng = np.random.default_rng(4)
    qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, E))
    qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, E)), Q_LO[:, None], Q_HI[:, None])
    a, b = torch.as_tensor(qa, device=dev), torch.as_tensor(qb, device=dev)
    t_e = timeit(lambda: engine.edge_feasibility(a, b, W, 5.0, mode="rne"), max(K, 5))
    ff = engine.edge_feasibility(a, b, W, 5.0, mode="rne")
    evaluated = int(torch.clamp(((ff + 32) // 32) * 32, max=W).sum().item())   # waypoints actually evaluated (32/round)
    out["edges"] = {"workload": "configs[3]: 100k edges x 64 min-jerk waypoints, rne, 5 kg", "edges_per_s": E / t_e,
                    "waypoint_states_per_s_nominal": E * W / t_e, "waypoint_states_evaluated": evaluated,
                    "feasible_fraction": float((ff == W).float().mean().item()), "ms": t_e * 1e3}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 500_000
    reps_per_step = 1
    for _ in range(max(args.warmup, 1)):
        cpu_baseline_run(50_000, 1)
    t0 = time.perf_counter()
    total = 0
    cores = 1
    for _ in range(args.steps):
        _, cores, _ = cpu_baseline_run(n_sample, reps_per_step)
        total += n_sample * reps_per_step
    dt = time.perf_counter() - t0
    value = total / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d-state seeded subsample of the 1M-state workload per step, C oracle port "
                                   "of rne.py" % n_sample,
                         "python_port_states_per_s_per_core": python_port_rate(),
                         "python_port_note": "oracle/rne_numpy_port.py: NumPy restatement at rne.py's own per-call "
                                             "granularity, 300 states on one core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the feasibility masks are all-gathered each step: p2p = peer stores fused into "
                         "the torque kernel (tcmp_rne_batch_scatter); nccl = all_gather_into_tensor on a side stream")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from torque_constrained_motion_planning_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    W = max(args.warmup, 3)
    K = args.steps

    # ---- synthetic inputs, resident in HBM before the timed region -------------------------------
    sets = []
    for s in range(N_SETS):
        q, qd, qdd, mass = sample_states(N_STATES, seed=2 + 1000 * rank + s)
        sets.append(tuple(torch.as_tensor(a, device=dev) for a in (q, qd, qdd, mass)))
    host0 = sample_states(N_STATES, seed=2 + 1000 * rank)
    from torque_constrained_motion_planning_b200.distributed import OverlappedGather, PeerMaskBuffer
    gather = None
    peer = None
    if world > 1 and args.gather == "nccl":
        gather = OverlappedGather((N_STATES,), torch.uint8, dev)
    elif world > 1:
        try:
            peer = PeerMaskBuffer(N_STATES)   # gathered [world][N_STATES] mask buffer, written by every rank's kernel
            ok_flag = torch.ones(1, device=dev)
        except Exception as e:                # CUDA IPC unavailable in this container: every rank must agree
            print("rank %d: peer-store gather unavailable (%s)" % (rank, e), file=sys.stderr)
            ok_flag = torch.zeros(1, device=dev)
        dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
        if ok_flag.item() < 1:
            if peer is not None:
                peer.close()
            peer = None
            args.gather = "nccl (p2p unavailable)"
            gather = OverlappedGather((N_STATES,), torch.uint8, dev)

    # outputs are preallocated once: a step is the kernel, not the caching allocator
    out_tau = torch.empty((7, N_STATES), dtype=torch.float64, device=dev)
    out_mask = torch.empty((N_STATES,), dtype=torch.uint8, device=dev)

    def step(i, mode="rne"):
        q, qd, qdd, mass = sets[i % N_SETS]
        if peer is not None:
            # fused compute + all-gather: the kernel stores each mask byte into every rank's gathered buffer
            return peer.torque_test(q, qd, qdd, mass, mode=mode, out_tau=out_tau), None
        tau, ok = engine.torque_test_batch(q, qd, qdd, mass, mode=mode, out_tau=out_tau, out_mask=out_mask)
        if gather is not None:
            gather.submit(ok)      # NCCL all-gather of this step's mask on a side stream (overlaps step i+1)
        return tau, ok

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        if gather is not None:
            gather.join()          # the timed region ends when the last mask all-gather has landed
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()              # covers warm-up, the timed region and the sustained hold below
    for i in range(W):
        step(i)
    ms_total = timed(step, K)        # the contract's number: W warm-up steps, then exactly K timed steps
    # K steps last ~1 ms, far below nvidia-smi's sampling period: afterwards hold the same step back to back for
    # ~1.5 s so the clock / throttle record reflects load, and report that rate too (`sustained`).
    # the step count comes from ms_total (already max-reduced, so identical on every rank: no rank may issue
    # more collectives than another)
    held = max(200, int(1.5 / (ms_total / K * 1e-3)) // 200 * 200)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for h in range(0, held, 200):
        for i in range(200):
            step(h + i)
        torch.cuda.synchronize()
    if gather is not None:
        gather.join()
    e1.record()
    torch.cuda.synchronize()
    sustained = world * N_STATES * held / (e0.elapsed_time(e1) * 1e-3)
    # kernel-only duration (no gather epilogue) for the roofline of the dominant kernel, same stream / events.
    # At N > 1 it is averaged over >= 200 launches: the K-step region is ~1 ms, of which the barrier-release skew
    # between ranks (max over ranks is reported) would be a large part.
    k_kernel = K if world == 1 else max(K, 200)
    ms_kernel = timed(lambda i: engine.torque_test_batch(*sets[i % N_SETS], mode="rne", out_tau=out_tau,
                                                         out_mask=out_mask), k_kernel) if world > 1 else ms_total
    clocks = sampler.stop() if rank == 0 else None
    value = world * N_STATES * K / (ms_total * 1e-3)
    kernel_s = ms_kernel * 1e-3 / k_kernel

    modes = {}
    for mode in ("nov", "dyn"):
        for i in range(3):
            step(i, mode)
        ms = timed(lambda i, m=mode: step(i, m), K)
        modes[mode] = world * N_STATES * K / (ms * 1e-3)
    modes["rne"] = value
    if world == 1:   # the optional fp32 path (1e-4 relative), same launch geometry
        f32 = [tuple(t.float() for t in s_) for s_ in sets[:2]]
        for i in range(3):
            engine.torque_test_batch(*f32[i % 2], mode="rne", dtype="f32")
        ms = timed(lambda i: engine.torque_test_batch(*f32[i % 2], mode="rne", dtype="f32"), K)
        modes["rne_f32"] = N_STATES * K / (ms * 1e-3)
        del f32
    # mask-only rne (the planner's actual need: 177 B/state)
    ms = timed(lambda i: engine.torque_test_batch(*sets[i % N_SETS], mode="rne", want_tau=False, out_mask=out_mask), K)
    modes["rne_mask_only"] = world * N_STATES * K / (ms * 1e-3)
    if peer is not None:
        # correctness of the fused gather: every rank must now hold every rank's mask of the last step
        last = (K - 1) % N_SETS
        step(last)
        peer.barrier()
        mine = engine.torque_test_batch(*sets[last], mode="rne", want_tau=False)[1]
        ref = torch.empty((world, N_STATES), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref, mine)
        assert torch.equal(ref, peer.gathered), "peer-store gather != NCCL all-gather"

    # ---- the other hot-path workloads (BASELINE.json configs[2], configs[3]); single-GPU runs only -------
    extras = {}
    if world == 1:
        extras = run_extras(engine, dev, K)

    # ---- end to end through the host-buffer C-ABI call (pinned host arrays, H2D + D2H inside) -------
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
    hq, hqd, hqdd, hm = (pin(a) for a in host0)
    htau = torch.empty((7, N_STATES), dtype=torch.float64).pin_memory()
    hok = torch.empty((N_STATES,), dtype=torch.uint8).pin_memory()
    ws = engine.Workspace(chunk_states=1 << 18)
    nq, nqd, nqdd, nm, ntau, nok = (t.numpy() for t in (hq, hqd, hqdd, hm, htau, hok))

    def e2e_step(_i):
        engine.torque_test_batch_host_into(ws, "rne", "f64", nq, nqd, nqdd, nm, 0.0, 0.01, ntau, nok)

    for i in range(3):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * N_STATES * K / e2e_s

    # sanity: the host path and the device path agree bit for bit on the same inputs
    tau_d, ok_d = engine.torque_test_batch(*sets[0], mode="rne")
    assert torch.equal(tau_d.cpu(), htau) and torch.equal(ok_d.cpu(), hok)

    # what bounds it: the same call without the torque read-back (the planner's predicate needs the mask only), and
    # a plain pinned cudaMemcpy of one input array in each direction (the PCIe ceiling of this box, this run)
    def timed_host(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t_ = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t_) / reps
    mask_only_s = timed_host(lambda: engine.torque_test_batch_host_into(ws, "rne", "f64", nq, nqd, nqdd, nm, 0.0, 0.01,
                                                                       None, nok), max(3, K // 2))
    dq = torch.empty((7, N_STATES), dtype=torch.float64, device=dev)
    h2d_s = timed_host(lambda: dq.copy_(hq, non_blocking=True), 10)
    d2h_s = timed_host(lambda: htau.copy_(dq, non_blocking=True), 10)
    link = {"h2d_gbs_in_call": N_STATES * 176 / (e2e_s / K) / 1e9, "h2d_gbs_in_call_mask_only": N_STATES * 176 / mask_only_s / 1e9,
            "h2d_gbs_plain_memcpy": N_STATES * 56 / h2d_s / 1e9, "d2h_gbs_plain_memcpy": N_STATES * 56 / d2h_s / 1e9,
            "mask_only_states_per_s": N_STATES / mask_only_s,
            "note": "per rank; the call moves 176 B/state host->device, so PCIe bounds it at plain-memcpy GB/s / 176 B"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fp64_peak = max(engine.fp64_peak(2048) for _ in range(3))
        achieved_tf = FLOPS_PER_STATE * N_STATES / kernel_s / 1e12
        achieved_gbs = BYTES_PER_STATE * N_STATES / kernel_s / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved_tf / (fp64_peak / 1e12),
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, from the ncu --set full
                # capture in profiles/r01/ncu_full_rne_batch_kernel_final_raw.csv (168.0 MB read + 24.9 MB written at
                # kernel end; the remaining dirty lines are still in the 126 MB L2) -- <= 233 MB algorithmic
                "traffic": 192.9e6,
                "peak_source": "tcmp_fp64_peak DFMA microbenchmark measured in this run (MEASURED_PEAKS.json "
                               "carries no FP64 entry; datasheet 37.2 TFLOP/s)",
                "flops_per_state": FLOPS_PER_STATE, "kernel": "rne_batch_kernel<double,DYN,!TOOL,tau,mask>",
                # what the kernel EXECUTES per state (static SASS count of the loop body, scripts/sass_mix.sh):
                # 419 DFMA + 180 DMUL + 70 DADD; pipe_frac = FP64 instructions issued / the DFMA rate behind `peak`
                "note": "achieved / frac use SURVEY 8d's ALGORITHMIC 1654 FLOP per state; the kernel executes 1088 "
                        "(customised recursion, regrouped parameters, table-driven sincos), so frac can reach 1 "
                        "while the FP64 pipe is at executed.pipe_frac",
                "executed": {"fp64_instr_per_state": FP64_INSTR_PER_STATE, "flops_per_state": EXEC_FLOPS_PER_STATE,
                             "achieved": EXEC_FLOPS_PER_STATE * N_STATES / kernel_s / 1e12,
                             "pipe_frac": FP64_INSTR_PER_STATE * N_STATES / kernel_s / (fp64_peak / 2.0)},
                "kernel_ms": kernel_s * 1e3, "kernel_launches_averaged": k_kernel,
                "hbm": {"achieved": achieved_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": achieved_gbs / peaks.get("hbm_gbs"), "bytes_per_state": BYTES_PER_STATE,
                        "peak_source": peak_src},
            },
            "modes": modes,
            "gather": "none (N = 1)" if world == 1 else args.gather,
            "extras": extras,
            "sustained": {"value": sustained, "unit": UNIT, "steps": held,
                          "note": "same step held back to back for >= 1.5 s (device-timed); clocks sampled over it"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(N_STATES * 176), "d2h_bytes_per_step": int(N_STATES * 57),
                    "api": "tcmp_rne_batch_host (pinned host SoA arrays, 3-stage chunked H2D/kernel/D2H pipeline)",
                    "link": link},
            "gpu_launches": K,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            v, cores, dt = cpu_baseline_run(N_STATES, 3)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "3 passes over the same 1M-state workload (%.1f s wall), C oracle port of "
                                             "rne.py with OpenMP" % dt,
                                   "python_port_states_per_s_per_core": python_port_rate(),
                                   "python_port_note": "oracle/rne_numpy_port.py: NumPy restatement at rne.py's own "
                                                       "per-call granularity (np.block / np.linalg.inv per link), "
                                                       "300 states on one core"}
            if world == 1:
                out["cpu_baseline"]["other_workloads"] = cpu_baseline_extras()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
