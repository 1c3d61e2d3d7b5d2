"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), ``torch.distributed`` for plumbing.

The units of every BASELINE config are independent (states, edges, (pose, free) solves), so they
shard as contiguous equal blocks ``[r*n/G, (r+1)*n/G)`` per rank with NO collective on the data path.
The only exchange is afterwards and tiny: an all-gather of the feasibility masks / first-failure
indices (1 B or 4 B per unit) and, for IK, of the solution counts and solution sets -- NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests.  The per-rank compute function is injectable so the
host logic (bounds, padding, gather order) is testable on CPU with world_size 2.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of rank ``rank``: [lo, hi).  Blocks differ in size by at most one unit."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_ragged(local, n_total: int, group=None):
    """Gather per-rank blocks (first dim = units of this rank, sizes from shard_bounds) into the full
    array on every rank.  Blocks are padded to the largest block so a single all_gather_into_tensor
    (NCCL: one NVSwitch collective) moves them."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad_shape = (mx,) + tuple(local.shape[1:])
    padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world,) + pad_shape, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * mx, *local.shape[1:]), padded, group=group)
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def sharded_torque_test(q, qd=None, qdd=None, payload_mass=0.0, mode="rne", group=None, gather=True,
                        compute: Optional[Callable] = None, **kw):
    """Every rank holds (or can index) the full SoA arrays ``[7][n]``; it evaluates only its block and, if
    ``gather``, all ranks end up with the full feasibility mask ``[n]``.  Returns (mask, (lo, hi))."""
    import torch
    if compute is None:
        from . import engine
        compute = engine.torque_test_batch
    rank, world = _world(group)
    n = int(q.shape[1])
    lo, hi = shard_bounds(n, rank, world)
    sl = lambda a: None if a is None else a[:, lo:hi].contiguous() if hasattr(a, "contiguous") else np.ascontiguousarray(a[:, lo:hi])
    pm = payload_mass if np.ndim(payload_mass) == 0 else payload_mass[lo:hi]
    _, ok = compute(sl(q), sl(qd), sl(qdd), pm, mode=mode, want_tau=False, **kw)
    if not gather or world == 1:
        return ok, (lo, hi)
    if not torch.is_tensor(ok):
        ok = torch.as_tensor(ok)
    return all_gather_ragged(ok, n, group), (lo, hi)


def sharded_edge_feasibility(qa, qb, n_waypoints=64, payload_mass=0.0, mode="rne", group=None, gather=True,
                             compute: Optional[Callable] = None, **kw):
    """Edges shard by contiguous blocks; an edge never straddles ranks, so the per-edge first-failure
    reduction needs no cross-GPU step.  Returns (first_fail [n_edges] on every rank, (lo, hi))."""
    import torch
    if compute is None:
        from . import engine
        compute = engine.edge_feasibility
    rank, world = _world(group)
    n = int(qa.shape[1])
    lo, hi = shard_bounds(n, rank, world)
    cut = lambda a: a[:, lo:hi].contiguous() if hasattr(a, "contiguous") else np.ascontiguousarray(a[:, lo:hi])
    ff = compute(cut(qa), cut(qb), n_waypoints, payload_mass, mode=mode, **kw)
    if not gather or world == 1:
        return ff, (lo, hi)
    if not torch.is_tensor(ff):
        ff = torch.as_tensor(ff)
    return all_gather_ragged(ff, n, group), (lo, hi)


def sharded_ik(rot9, trans3, free, group=None, gather=True, compute: Optional[Callable] = None):
    """Poses shard by contiguous blocks (each pose keeps its whole free-joint sweep on one rank).
    Returns (sols [n*n_free][8][7], counts [n*n_free], (lo, hi)) gathered on every rank."""
    import torch
    if compute is None:
        from . import engine
        compute = lambda r, t, f: engine.ik_batch(r, t, f)[:2]
    rank, world = _world(group)
    n = int(rot9.shape[1])
    lo, hi = shard_bounds(n, rank, world)
    cut = lambda a: a[:, lo:hi].contiguous() if hasattr(a, "contiguous") else np.ascontiguousarray(a[:, lo:hi])
    bcast = len(free.shape) == 1
    n_free = int(free.shape[0])
    sols, counts = compute(cut(rot9), cut(trans3), free if bcast else cut(free))
    if not gather or world == 1:
        return sols, counts, (lo, hi)
    if not torch.is_tensor(sols):
        sols, counts = torch.as_tensor(sols), torch.as_tensor(counts)
    m = hi - lo
    sols_g = all_gather_ragged(sols.reshape(m, n_free * 56), n, group).reshape(n * n_free, 8, 7)
    counts_g = all_gather_ragged(counts.reshape(m, n_free), n, group).reshape(n * n_free)
    return sols_g, counts_g, (lo, hi)


class OverlappedGather:
    """All-gather of per-rank result blocks (masks, first-failure indices) on a side stream so the
    collective of step i overlaps the kernel of step i+1.  The payload is tiny (1 B/state), i.e. the
    collective is pure latency (~20 us on NVSwitch) against a ~60 us kernel -- serialising them costs a
    quarter of the step, overlapping them hides it.  ``submit`` may be called once per step; ``join`` makes
    the current stream wait for everything submitted (call it before reading ``out`` or stopping a timer)."""

    def __init__(self, block_shape, dtype, device, group=None):
        import torch
        self._torch = torch
        self.group = group
        _, world = _world(group)
        self.world = world
        self.stream = torch.cuda.Stream(device=device)
        self.out = torch.empty((world,) + tuple(block_shape), dtype=dtype, device=device)

    def submit(self, local):
        torch = self._torch
        if self.world == 1:
            self.out[0].copy_(local, non_blocking=True)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            _dist().all_gather_into_tensor(self.out.view(-1, *self.out.shape[2:]), local, group=self.group)
        local.record_stream(self.stream)

    def join(self):
        self._torch.cuda.current_stream().wait_stream(self.stream)
        return self.out


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view memory owned by libtcmp.so."""

    def __init__(self, ptr, n, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerMaskBuffer:
    """Gathered feasibility-mask buffer ``[SLOTS][world][n_per_rank]`` (``gathered`` = the copy of the latest step,
    ``[world][n_per_rank]``) that every rank's torque kernel writes into
    DIRECTLY (tcmp_rne_batch_scatter): the all-gather of the masks fused into the producing kernel as a
    peer-store epilogue over NVLink/NVSwitch.  Each rank allocates its copy with tcmp_peer_alloc (cudaMalloc +
    CUDA IPC), the 64-byte handles are exchanged once through torch.distributed, and peers are mapped with
    tcmp_peer_open.  After ``torque_test`` + a stream sync + ``barrier()`` every rank holds every mask.

    Completion and ordering, without a host barrier or a collective call (``overlap_gather=True``).  After every
    scatter kernel a side stream runs ``signal()`` (tcmp_peer_signal: this rank publishes the step's epoch into every
    rank's sync block) and ``wait()`` (tcmp_peer_wait: holds the side stream until every rank has published it); both
    ride under the NEXT step's kernel on the producer stream, so the gather's completion costs the step nothing.  The
    buffer cycles through ``SLOTS`` = 3 copies and the scatter kernel of step i + 3 is ordered behind this rank's
    wait(i + 1).  Contract for a consumer of step i's masks: enqueue the reads on ``side`` right after the step (they
    then run after wait(i) and before signal(i + 1)); a peer can only overwrite copy i % 3 in step i + 3, i.e. after it
    has seen every rank publish i + 1, which this rank does after those reads.  ``join()`` makes the current stream
    wait for the last wait.  Everything is plain kernel launches and events: CUDA-graph capturable (call ``reset()``
    before a capture so no event recorded outside the capture is waited on inside).

    ``barrier()`` (stream sync + torch.distributed barrier) remains for callers that synchronise on the host: with one
    barrier per step, step i + 1's writes go to another copy than step i's, and a copy is written again only three
    steps later, so readers of step i have until they enter the barrier of step i + 2.  (With a single
    copy a fast rank's next step could overwrite a row a slower rank was still reading: ADVICE r01.)"""

    itemsize = 1
    typestr = "|u1"
    SLOTS = 3

    def __init__(self, n_per_rank: int, group=None, multicast=False):
        import ctypes

        import torch
        from . import _lib
        self._lib = _lib.load()
        self._check = _lib.check
        self.group = group
        self.rank, self.world = _world(group)
        if self.world > 8:
            raise ValueError("at most 8 peers (one NVSwitch node)")
        self.n = int(n_per_rank)
        nbytes = self.SLOTS * self.n * self.world * self.itemsize
        self._step = 0
        self._own = ctypes.c_void_p()
        self._peers = []
        self._symm = None          # (tensor, handle) when the buffer lives in torch symmetric memory
        self.mc_ptr = None         # NVSwitch multicast address of the gathered buffer (NVLS), or None
        dev = torch.device("cuda", torch.cuda.current_device())
        ptrs = (ctypes.c_void_p * self.world)()
        if self.world > 1 and multicast:
            # NVLS (opt-in: measured 1 % faster than unicast peer stores at N = 8, and it depends on torch's symmetric
            # memory rendezvous): allocate the buffer in symmetric memory (torch does the cuMem / cuMulticast plumbing and the file-
            # descriptor exchange) -- unicast peer pointers AND one multicast address for the same physical pages.
            # Every rank must end up on the same path, so the outcome is agreed with an all-reduce.
            ok = 0
            try:
                import torch.distributed._symmetric_memory as symm
                t = symm.empty(nbytes, dtype=torch.uint8, device=dev)
                name = (group if group is not None else _dist().group.WORLD).group_name
                h = symm.rendezvous(t, group=name)
                ok = int(h.multicast_ptr != 0 and len(h.buffer_ptrs) == self.world)
            except Exception:
                ok = 0
            flag = torch.tensor([ok], device=dev)
            _dist().all_reduce(flag, op=_dist().ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                t.zero_()
                self._symm = (t, h)
                self.mc_ptr = int(h.multicast_ptr)
                for r in range(self.world):
                    ptrs[r] = int(h.buffer_ptrs[r])
        if self._symm is None:
            handle = ctypes.create_string_buffer(64)
            self._check(self._lib.tcmp_peer_alloc(ctypes.byref(self._own), nbytes, handle))
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=dev)
            if self.world > 1:
                allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
                _dist().all_gather_into_tensor(allh, mine, group=group)
                allh = allh.cpu().numpy()
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = self._own.value
                else:
                    p = ctypes.c_void_p()
                    self._check(self._lib.tcmp_peer_open(bytes(allh[r].tobytes()), ctypes.byref(p)))
                    self._peers.append(p)
                    ptrs[r] = p.value
        self._ptrs = ptrs
        # completion flags of the fused gather: one TCMP_PEER_SYNC_BYTES block per rank, peers mapped like the buffers
        self._sync_own = ctypes.c_void_p()
        shandle = ctypes.create_string_buffer(64)
        self._check(self._lib.tcmp_peer_alloc(ctypes.byref(self._sync_own), 128, shandle))
        smine = torch.tensor(list(shandle.raw), dtype=torch.uint8, device=dev)
        sptrs = (ctypes.c_void_p * self.world)()
        if self.world > 1:
            sall = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
            _dist().all_gather_into_tensor(sall, smine, group=group)
            sall = sall.cpu().numpy()
        for r in range(self.world):
            if r == self.rank:
                sptrs[r] = self._sync_own.value
            else:
                p = ctypes.c_void_p()
                self._check(self._lib.tcmp_peer_open(bytes(sall[r].tobytes()), ctypes.byref(p)))
                self._peers.append(p)
                sptrs[r] = p.value
        self._sync_ptrs = sptrs
        self.side = torch.cuda.Stream(device=dev)     # carries signal / wait (and a consumer's reads)
        self._done = {}                               # step -> event recorded on `side` after wait(step)
        if self._symm is not None:
            flat = self._symm[0].view(torch.uint8 if self.itemsize == 1 else torch.int32)
        else:
            flat = torch.as_tensor(_DevArray(self._own.value, self.SLOTS * self.n * self.world, self.typestr), device=dev)
        self._slots = flat.view(self.SLOTS, self.world, self.n)
        self.gathered = self._slots[0]
        if self.world > 1:
            torch.cuda.synchronize()
            _dist().barrier(group=group)      # nobody scatters before every rank's buffer is zeroed and mapped

    def _next_offset(self):
        """Element offset of this rank's row in the copy the coming step writes; ``gathered`` follows it."""
        slot = self._step % self.SLOTS
        self._step += 1
        self.gathered = self._slots[slot]
        return (slot * self.world + self.rank) * self.n

    def torque_test(self, q, qd=None, qdd=None, payload_mass=0.0, mode="rne", payload_threshold=0.01,
                    want_tau=True, out_tau=None, overlap_gather=False):
        """Evaluate this rank's block (q/qd/qdd [7][n] CUDA tensors) and store its mask into row ``rank`` of the
        gathered buffer on EVERY rank.  Returns tau [7][n] (or None).  ``overlap_gather=True``: completion flags on
        the side stream as described in the class docstring (then use ``join()``, not ``barrier()``)."""
        import torch
        from ._lib import DTYPE, MODE
        n = int(q.shape[1])
        assert n <= self.n
        scalar, pm = (float(payload_mass), None) if np.ndim(payload_mass) == 0 else (0.0, payload_mass)
        tau = (out_tau if out_tau is not None else torch.empty((7, n), dtype=torch.float64, device=q.device)) \
            if want_tau else None
        ptr = lambda t: None if t is None else int(t.data_ptr())
        if overlap_gather:
            self.before_step()
        if self.mc_ptr is not None:
            self._check(self._lib.tcmp_rne_batch_scatter_mc(
                MODE[mode], DTYPE["f64"], n, ptr(q), ptr(qd), ptr(qdd), ptr(pm), scalar, float(payload_threshold),
                ptr(tau), self.world, self._ptrs, self.mc_ptr, self._next_offset(),
                int(torch.cuda.current_stream().cuda_stream)))
        else:
            self._check(self._lib.tcmp_rne_batch_scatter(
                MODE[mode], DTYPE["f64"], n, ptr(q), ptr(qd), ptr(qdd), ptr(pm), scalar, float(payload_threshold),
                ptr(tau), self.world, self._ptrs, self._next_offset(), int(torch.cuda.current_stream().cuda_stream)))
        if overlap_gather:
            self._after_step()
        return tau

    def before_step(self):
        """The coming step (index self._step) overwrites the copy of step - SLOTS: order it behind wait(step - 2)."""
        import torch
        ev = self._done.get(self._step - (self.SLOTS - 1))
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _after_step(self):
        """signal + wait of the step just launched (self._step - 1) on the side stream."""
        import torch
        step = self._step - 1
        launched = torch.cuda.Event()
        launched.record(torch.cuda.current_stream())
        self.side.wait_event(launched)
        with torch.cuda.stream(self.side):
            self.signal()
            self.wait()
            done = torch.cuda.Event()
            done.record(self.side)
        self._done[step] = done
        self._done.pop(step - self.SLOTS, None)

    def check(self):
        """Synchronise and raise if a tcmp_peer_wait gave up on a rank (~20 s without its epoch: that rank is gone)."""
        import torch
        torch.cuda.synchronize()
        word = torch.as_tensor(_DevArray(self._sync_own.value, 128, "|u1"),
                               device=torch.device("cuda", torch.cuda.current_device())).view(torch.int32)[19]
        if int(word.item()) != 0:
            raise RuntimeError("tcmp_peer_wait timed out waiting for rank %d" % (int(word.item()) - 1))

    def join(self):
        """Current stream waits until the last step's gather is complete on this rank."""
        import torch
        if self._done:
            torch.cuda.current_stream().wait_event(self._done[max(self._done)])

    def reset(self):
        """Forget the recorded completion events (host side only; synchronise first).  Call before capturing steps into
        a CUDA graph and after the capture: events of one regime must not be waited on in the other."""
        self._done = {}

    def push(self, local):
        """The unfused form of a step: ``local`` (this rank's result block, a CUDA tensor of n * itemsize bytes produced
        on the current stream) is copied into row ``rank`` of every rank's gathered buffer by tcmp_peer_push on the
        side stream, followed by signal / wait -- all under the next step's kernel.  The caller alternates ``local``
        between two buffers and calls ``before_step()`` ahead of each producing kernel (it orders the kernel behind
        the wait of two steps ago, by when the push that read this buffer has finished)."""
        import torch
        off = self._next_offset() * self.itemsize
        launched = torch.cuda.Event()
        launched.record(torch.cuda.current_stream())
        self.side.wait_event(launched)
        with torch.cuda.stream(self.side):
            self._check(self._lib.tcmp_peer_push(int(local.data_ptr()), int(local.numel() * local.element_size()),
                                                 self.world, self._ptrs, off, int(self.side.cuda_stream)))
            self.signal()
            self.wait()
            done = torch.cuda.Event()
            done.record(self.side)
        step = self._step - 1
        self._done[step] = done
        self._done.pop(step - self.SLOTS, None)

    def signal(self):
        """Publish this rank's completion after a scatter kernel that does not signal itself (tcmp_peer_signal)."""
        import torch
        self._check(self._lib.tcmp_peer_signal(self.rank, self.world, self._sync_ptrs,
                                               int(torch.cuda.current_stream().cuda_stream)))

    def wait(self):
        """Device-side wait (tcmp_peer_wait, one tiny kernel on the current stream): returns at once on the host; work
        enqueued after it runs when every rank has published this rank's current epoch, i.e. ``gathered`` is
        complete.  No host barrier, no collective call; CUDA-graph capturable."""
        import torch
        self._check(self._lib.tcmp_peer_wait(self._sync_own, self.world, int(torch.cuda.current_stream().cuda_stream)))

    def barrier(self):
        """Order every rank's peer stores before anyone reads ``gathered``."""
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            _dist().barrier(group=self.group)

    def close(self):
        for p in self._peers:
            self._lib.tcmp_peer_close(p)
        self._peers = []
        if self._sync_own:
            self._lib.tcmp_peer_free(self._sync_own)
            self._sync_own = None
        self.gathered = self._slots = None
        self._symm = None
        self.mc_ptr = None
        if self._own:
            self._lib.tcmp_peer_free(self._own)
            self._own = None


class PeerIndexBuffer(PeerMaskBuffer):
    """Gathered ``int32 [world][n_per_rank]`` buffer for per-edge first-failure indices, written by every rank's
    edge kernel directly (tcmp_edge_feasibility_scatter) -- configs[3] "sharded at 2/4/8 B200" without a
    collective call on the step."""

    itemsize = 4
    typestr = "<i4"

    def edge_feasibility(self, qa, qb, n_waypoints=64, payload_mass=0.0, mode="rne", payload_threshold=0.01,
                         static_only=False, overlap_gather=False):
        """Evaluate this rank's block of edges (qa/qb [7][n] CUDA fp64 tensors); row ``rank`` of ``gathered`` on
        every rank receives the first-failure indices."""
        import torch
        from ._lib import MODE
        n = int(qa.shape[1])
        assert n <= self.n
        if overlap_gather:
            self.before_step()
        self._check(self._lib.tcmp_edge_feasibility_scatter(
            MODE[mode], n, int(n_waypoints), int(qa.data_ptr()), int(qb.data_ptr()), float(payload_mass),
            float(payload_threshold), int(static_only), self.world, self._ptrs, self._next_offset(),
            int(torch.cuda.current_stream().cuda_stream)))
        if overlap_gather:
            self._after_step()
