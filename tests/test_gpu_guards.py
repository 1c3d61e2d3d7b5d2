"""GPU: own bounds checks (compute-sanitizer is closed on this pool): every output buffer is carved out of a
larger sentinel-filled allocation; after the call the guard bands on both sides must be untouched, at ragged
sizes that end mid-warp / mid-CTA."""
import numpy as np
import pytest

from conftest import Q_HI, Q_LO, sample_edges, sample_states

pytestmark = pytest.mark.gpu

GUARD = 4096


def guarded(torch, shape, dtype):
    n = int(np.prod(shape))
    sentinel = {torch.float64: -7.25e300, torch.float32: -7.25e30, torch.uint8: 0xA5, torch.int32: -1234567}[dtype]
    big = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device="cuda")
    return big, big[GUARD:GUARD + n].view(*shape), sentinel


def intact(big, n, sentinel):
    return bool((big[:GUARD] == sentinel).all()) and bool((big[GUARD + n:] == sentinel).all())


@pytest.mark.parametrize("n", [1, 31, 129, 1000, 75777])
def test_output_guard_bands(n):
    import torch
    from torque_constrained_motion_planning_b200 import _lib
    lib = _lib.load()
    st = int(torch.cuda.current_stream().cuda_stream)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    q, qd, qdd, m = (dev(a) for a in sample_states(n, seed=n))
    # tcmp_rne_batch
    bt, tau, s_t = guarded(torch, (7, n), torch.float64)
    bm, mask, s_m = guarded(torch, (n,), torch.uint8)
    for mode in range(4):
        _lib.check(lib.tcmp_rne_batch(mode, 0, n, q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), m.data_ptr(), 0.0, 0.01,
                                      tau.data_ptr(), mask.data_ptr(), st))
    torch.cuda.synchronize()
    assert intact(bt, 7 * n, s_t) and intact(bm, n, s_m)
    assert bool((mask <= 1).all()) and bool(torch.isfinite(tau).all())
    # tcmp_edge_feasibility
    qa, qb = (dev(a) for a in sample_edges(n, seed=n))
    bf, ff, s_f = guarded(torch, (n,), torch.int32)
    for W in (1, 33, 64):
        _lib.check(lib.tcmp_edge_feasibility(0, 0, n, W, qa.data_ptr(), qb.data_ptr(), 5.0, 0.01, 0, ff.data_ptr(), st))
        torch.cuda.synchronize()
        assert intact(bf, n, s_f) and bool(((ff >= 0) & (ff <= W)).all())
    # tcmp_fk_batch / tcmp_ik_batch
    b1, trans, s1 = guarded(torch, (3, n), torch.float64)
    b2, rot, s2 = guarded(torch, (9, n), torch.float64)
    _lib.check(lib.tcmp_fk_batch(n, q.data_ptr(), trans.data_ptr(), rot.data_ptr(), st))
    nf = 3
    free = dev(np.random.default_rng(n).uniform(Q_LO[6], Q_HI[6], size=(nf, n)))
    b3, sols, s3 = guarded(torch, (n * nf, 8, 7), torch.float64)
    b4, cnt, s4 = guarded(torch, (n * nf,), torch.int32)
    b5, stat, s5 = guarded(torch, (n * nf,), torch.uint8)
    _lib.check(lib.tcmp_ik_batch(n, rot.data_ptr(), trans.data_ptr(), free.data_ptr(), nf, 0, sols.data_ptr(),
                                 cnt.data_ptr(), stat.data_ptr(), st))
    torch.cuda.synchronize()
    assert intact(b1, 3 * n, s1) and intact(b2, 9 * n, s2)
    assert intact(b3, n * nf * 56, s3) and intact(b4, n * nf, s4) and intact(b5, n * nf, s5)
    assert bool(((cnt >= 0) & (cnt <= 8)).all())
    assert bool((sols != s3).all())                      # every slot written (solutions or zero fill)
    # tcmp_traj_feasibility
    from torque_constrained_motion_planning_b200 import min_jerk_v2
    pts = np.random.default_rng(n).uniform(Q_LO, Q_HI, size=(4, 7))
    coeffs = dev(min_jerk_v2.coefficients_for_kernel(min_jerk_v2.minjerk_coefficients(pts)))
    S = max(1, n // 3)
    ns = 3 * S
    outs = [guarded(torch, (7, ns), torch.float64) for _ in range(4)]
    bk, msk, sk = guarded(torch, (ns,), torch.uint8)
    first = torch.full((1,), ns, dtype=torch.int32, device="cuda")
    _lib.check(lib.tcmp_traj_feasibility(0, 0, 3, S, coeffs.data_ptr(), 5.0, 0.01, outs[0][1].data_ptr(),
                                         outs[1][1].data_ptr(), outs[2][1].data_ptr(), outs[3][1].data_ptr(),
                                         msk.data_ptr(), first.data_ptr(), st))
    torch.cuda.synchronize()
    assert all(intact(b, 7 * ns, s) for b, _, s in outs) and intact(bk, ns, sk)
    assert 0 <= int(first.item()) <= ns


@pytest.mark.parametrize("n", [1, 33, 1000, 20011])
def test_round2_kernels_guard_bands(n):
    """The kernels added in round 2, same own-bounds-check scheme: the IK redo pass (every pose at the elbow
    singularity, so the second kernel writes all solution sets), the BASE-mode trajectory, the scatter kernel with a
    destination offset, tcmp_peer_push, and the completion-flag kernels on a sync block."""
    import ctypes
    import torch
    from torque_constrained_motion_planning_b200 import _lib, min_jerk_v2
    lib = _lib.load()
    st = int(torch.cuda.current_stream().cuda_stream)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    qh, qdh, qddh, mh = sample_states(n, seed=100 + n)
    qh[3] = 2.63084142381503
    qh[3, ::3] = 0.0
    q, qd, qdd, m = (dev(a) for a in (qh, qdh, qddh, mh))
    b1, trans, s1 = guarded(torch, (3, n), torch.float64)
    b2, rot, s2 = guarded(torch, (9, n), torch.float64)
    _lib.check(lib.tcmp_fk_batch(n, q.data_ptr(), trans.data_ptr(), rot.data_ptr(), st))
    nf = 2
    free = dev(np.vstack([qh[6:7], np.zeros((1, n))]))
    b3, sols, s3 = guarded(torch, (n * nf, 8, 7), torch.float64)
    b4, cnt, s4 = guarded(torch, (n * nf,), torch.int32)
    b5, stat, s5 = guarded(torch, (n * nf,), torch.uint8)
    _lib.check(lib.tcmp_ik_batch(n, rot.data_ptr(), trans.data_ptr(), free.data_ptr(), nf, 0, sols.data_ptr(),
                                 cnt.data_ptr(), stat.data_ptr(), st))
    torch.cuda.synchronize()
    assert intact(b3, n * nf * 56, s3) and intact(b4, n * nf, s4) and intact(b5, n * nf, s5)
    assert bool(((cnt >= 0) & (cnt <= 8)).all()) and bool((sols != s3).all()) and bool((stat < 16).all())
    assert int((stat & 1).sum()) >= n // 2            # the elbow branch (and with it the redo kernel) was taken
    # BASE-mode trajectory: samples and torques written, mask all 1
    pts = np.random.default_rng(n).uniform(Q_LO, Q_HI, size=(4, 7))
    coeffs = dev(min_jerk_v2.coefficients_for_kernel(min_jerk_v2.minjerk_coefficients(pts)))
    S = max(1, n // 3)
    ns = 3 * S
    outs = [guarded(torch, (7, ns), torch.float64) for _ in range(4)]
    bk, msk, sk = guarded(torch, (ns,), torch.uint8)
    first = torch.full((1,), ns, dtype=torch.int32, device="cuda")
    _lib.check(lib.tcmp_traj_feasibility(3, 0, 3, S, coeffs.data_ptr(), 5.0, 0.01, outs[0][1].data_ptr(),
                                         outs[1][1].data_ptr(), outs[2][1].data_ptr(), outs[3][1].data_ptr(),
                                         msk.data_ptr(), first.data_ptr(), st))
    torch.cuda.synchronize()
    assert all(intact(b, 7 * ns, s) for b, _, s in outs) and intact(bk, ns, sk)
    assert all(bool((o != s).all()) for _, o, s in outs) and bool((msk == 1).all()) and int(first.item()) == ns
    # scatter into row 1 of a [3][n] gathered buffer, push into row 2; rows 0 and the guard bands stay untouched
    bg, gathered, sg = guarded(torch, (3, n), torch.uint8)
    ptrs = (ctypes.c_void_p * 1)(gathered.data_ptr())
    bt, tau, s_t = guarded(torch, (7, n), torch.float64)
    _lib.check(lib.tcmp_rne_batch_scatter(0, 0, n, q.data_ptr(), qd.data_ptr(), qdd.data_ptr(), m.data_ptr(), 0.0, 0.01,
                                          tau.data_ptr(), 1, ptrs, n, st))
    _lib.check(lib.tcmp_peer_push(gathered[1].data_ptr(), n, 1, ptrs, 2 * n, st))
    sync = torch.zeros(128, dtype=torch.uint8, device="cuda")
    sptrs = (ctypes.c_void_p * 1)(sync.data_ptr())
    _lib.check(lib.tcmp_peer_signal(0, 1, sptrs, st))
    _lib.check(lib.tcmp_peer_wait(sync.data_ptr(), 1, st))
    torch.cuda.synchronize()
    assert intact(bg, 3 * n, sg) and intact(bt, 7 * n, s_t) and bool((gathered[0] == sg).all())
    assert bool((gathered[1] <= 1).all()) and torch.equal(gathered[1], gathered[2])
    assert int(sync.view(torch.int64)[0]) == 1 and int(sync.view(torch.int64)[8]) == 1   # arrived[0] = epoch = 1


def test_concurrent_host_calls_share_the_default_workspace():
    """Python threads calling the host-array entry points at once: the shared default workspace is serialised by
    its lock, device calls on per-thread streams need none.  Every thread must get its own inputs' results."""
    import threading

    import oracle
    import torch
    from torque_constrained_motion_planning_b200 import engine as eng
    jobs = [sample_states(30_000 + 1000 * i, seed=50 + i) for i in range(6)]
    want = [oracle.torque_test_batch("rne", *j) for j in jobs]
    got = [None] * len(jobs)

    def host_job(i):
        got[i] = eng.torque_test_batch(*jobs[i], mode="rne")

    def device_job(i):
        with torch.cuda.stream(torch.cuda.Stream()):
            dev = [torch.as_tensor(a, device="cuda") for a in jobs[i]]
            tau, ok = eng.torque_test_batch(*dev, mode="rne")
            got[i] = (tau.cpu().numpy(), ok.cpu().numpy())

    threads = [threading.Thread(target=host_job if i % 2 == 0 else device_job, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for (tau, ok), (tau_o, ok_o) in zip(got, want):
        assert np.abs(tau - tau_o).max() < 1e-9 and np.array_equal(ok, ok_o)


def test_entry_points_are_cuda_graph_capturable():
    """The device entry points only enqueue (no allocation, no synchronisation, parameters by value), so a planner
    can capture its launch-bound inner loop once and replay it: K1 (with its programmatic-dependent-launch
    attribute), the static test, the edge kernel and FK are captured into one CUDA graph, the inputs are then
    overwritten IN PLACE, and each replay must reproduce what direct calls give on the new contents."""
    import torch
    from torque_constrained_motion_planning_b200 import engine as eng
    n, e = 20_000, 512
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    q, qd, qdd, m = (dev(a) for a in sample_states(n, seed=61))
    qa, qb = (dev(a) for a in sample_edges(e, seed=62))
    tau = torch.empty((7, n), dtype=torch.float64, device="cuda")
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    ok_nov = torch.empty(n, dtype=torch.uint8, device="cuda")
    lib = eng.load()
    ff = torch.empty(e, dtype=torch.int32, device="cuda")
    trans = torch.empty((3, n), dtype=torch.float64, device="cuda")
    rot = torch.empty((9, n), dtype=torch.float64, device="cuda")

    def enqueue():
        st = int(torch.cuda.current_stream().cuda_stream)
        eng.torque_test_batch(q, qd, qdd, m, mode="rne", out_tau=tau, out_mask=ok)
        eng.torque_test_batch(q, None, None, m, mode="nov", want_tau=False, out_mask=ok_nov)
        eng.check(lib.tcmp_edge_feasibility(0, 0, e, 64, qa.data_ptr(), qb.data_ptr(), 5.0, 0.01, 0, ff.data_ptr(), st))
        eng.check(lib.tcmp_fk_batch(n, q.data_ptr(), trans.data_ptr(), rot.data_ptr(), st))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        enqueue()                                   # warm-up outside capture (module load, occupancy queries)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue()
    for seed in (63, 64):
        nq, nqd, nqdd, nm = (dev(a) for a in sample_states(n, seed=seed))
        na, nb = (dev(a) for a in sample_edges(e, seed=seed + 10))
        for dst, src in ((q, nq), (qd, nqd), (qdd, nqdd), (m, nm), (qa, na), (qb, nb)):
            dst.copy_(src)
        for t in (tau, trans, rot):
            t.fill_(float("nan"))
        ok.fill_(7); ok_nov.fill_(7); ff.fill_(-9)
        graph.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in (tau, ok, ok_nov, ff, trans, rot)]
        want_tau, want_ok = eng.torque_test_batch(nq, nqd, nqdd, nm, mode="rne")
        _, want_nov = eng.torque_test_batch(nq, None, None, nm, mode="nov", want_tau=False)
        want_ff = eng.edge_feasibility(na, nb, 64, 5.0, mode="rne")
        want_trans, want_rot = eng.fk_batch(nq)
        for g, w in zip(got, (want_tau, want_ok, want_nov, want_ff, want_trans, want_rot)):
            assert torch.equal(g, w)
