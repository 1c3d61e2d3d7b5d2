#!/usr/bin/env python
"""oracle/make_golden.py -- generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these
fixtures are the pin for every oracle and CUDA parity test.  Sources executed:
  rne.py (rne, add_payload, remove_payload), min_jerk_v2.py (minjerk_coefficients,
  minjerk_trajectory), ikfast_panda_arm.cpp (ComputeIk / ComputeFk via oracle/_ref).
The torque-test closure bodies (panda_primitives.py:130-151,171-191) are restated in
ref_harness.ref_torque_test around the real rne.rne because the closures need PyBullet.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# panda_mod.urdf:127..283
Q_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
Q_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
V_LIM = np.array([2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61])


def sample_states(n, seed):
    """Config-2 distribution."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qd = rng.uniform(-V_LIM[:, None], V_LIM[:, None], size=(7, n))
    qdd = rng.uniform(-10.0, 10.0, size=(7, n))
    mass = rng.choice(np.array([0.0, 1.0, 3.0, 5.0]), size=n)
    return q, qd, qdd, mass


def sample_edges(n, seed):
    """Config-4 distribution."""
    rng = np.random.default_rng(seed)
    qa = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    qb = np.clip(qa + rng.normal(0.0, 0.5, size=(7, n)), Q_LO[:, None], Q_HI[:, None])
    return qa, qb


def gen_kats():
    qh = np.array(H.Q_HOME)
    z = np.zeros(7)
    q5 = np.array([0.5, 0.9, -0.3, -1.2, 0.7, 2.5, -1.0])
    cases = [
        (qh, z, z, 0.0), (qh, z, z, 1.0), (qh, z, z, 5.0),
        (qh, 0.1 * np.ones(7), 0.2 * np.ones(7), 1.0),
        (q5, np.array([1, -0.5, 0.8, -1.2, 0.3, 0.9, -2.0]), np.array([3, -4, 2, 5, -1, 2.5, -6.0]), 5.0),
        (z, z, z, 0.0),
        # payload rule edge cases: 0 < m <= 0.01 is NOT attached by the torque test (panda_primitives.py:178)
        (qh, z, z, 0.005), (qh, z, z, 0.01), (qh, z, z, 0.0100001),
    ]
    q = np.array([c[0] for c in cases]).T
    qd = np.array([c[1] for c in cases]).T
    qdd = np.array([c[2] for c in cases]).T
    m = np.array([c[3] for c in cases])
    tau = np.zeros((7, len(cases)))
    ok = np.zeros(len(cases), dtype=np.uint8)
    tau_raw = np.zeros((7, len(cases)))  # raw rne.rne with add_payload rule m > 0
    for i in range(len(cases)):
        o, t = H.ref_torque_test("rne", q[:, i], qd[:, i], qdd[:, i], m[i])
        tau[:, i], ok[i] = t, o
        tau_raw[:, i] = H.ref_rne(q[:, i], qd[:, i], qdd[:, i], m[i])
    np.savez(os.path.join(OUT, "kat_rne.npz"), q=q, qd=qd, qdd=qdd, mass=m, tau=tau, feasible=ok, tau_raw=tau_raw)
    print("kat_rne", ok.tolist())


def gen_states(n=3000, seed=2):
    q, qd, qdd, mass = sample_states(n, seed)
    tau_rne = np.zeros((7, n))
    ok_rne = np.zeros(n, dtype=np.uint8)
    tau_nov = np.zeros((7, n))
    ok_nov = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        o, t = H.ref_torque_test("rne", q[:, i], qd[:, i], qdd[:, i], mass[i])
        tau_rne[:, i], ok_rne[i] = t, o
        o, t = H.ref_torque_test("nov", q[:, i], None, None, mass[i])
        tau_nov[:, i], ok_nov[i] = t, o
    np.savez(os.path.join(OUT, "states_cfg2.npz"), seed=seed, q=q, qd=qd, qdd=qdd, mass=mass,
             tau_rne=tau_rne, feasible_rne=ok_rne, tau_nov=tau_nov, feasible_nov=ok_nov)
    print("states_cfg2: rne feasible %.3f nov feasible %.3f" % (ok_rne.mean(), ok_nov.mean()))


def custom_model(seed=21):
    """A deliberately different inertial set: every mass / COM / inertia perturbed, a massive link8 with an off-axis
    COM, a heavier hand with the real gripper's COM height, another payload lever and tighter limits."""
    rng = np.random.default_rng(seed)
    m = oracle.default_model()
    f = oracle.model_fields(m)
    f["mass"][:7] *= rng.uniform(0.7, 1.3, size=7)
    f["mass"][7] = 0.2
    f["mass"][8] = 1.1
    f["com"][:7] += rng.normal(0, 0.01, size=(7, 3))
    f["com"][7] = [0.01, -0.02, 0.03]
    f["com"][8] = [0.005, 0.01, 0.06]
    for k in range(9):
        A = rng.normal(0, 1, size=(3, 3))
        P = 0.002 * (A @ A.T)   # symmetric positive semi-definite perturbation
        ixx, ixy, ixz, iyy, iyz, izz = f["inertia"][k]
        I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]]) + P
        f["inertia"][k] = [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]
    f["payload_radius"][0] = 0.21
    f["tool_z"][0] = 0.13
    f["torque_limit"][:] = [80, 70, 60, 50, 11, 9, 12]
    return m


def gen_model(n=400, seed=22):
    """Pins the model-parametrised form of the oracle: rne.rne executed with the module's inertial lists overwritten."""
    model = custom_model()
    f = oracle.model_fields(model)
    q, qd, qdd, mass = sample_states(n, seed)
    tau = np.zeros((7, n))
    tau_static = np.zeros((7, n))
    z = np.zeros(7)
    with H.ref_inertial_override(f):
        for i in range(n):
            mp = mass[i] if mass[i] > 0.01 else 0.0
            tau[:, i] = H.ref_rne_payload_radius(q[:, i], qd[:, i], qdd[:, i], mp, f["payload_radius"][0])
            tau_static[:, i] = H.ref_rne_payload_radius(q[:, i], z, z, mp, f["payload_radius"][0])
    # the tables are back: the stock KAT must reproduce
    assert np.allclose(H.ref_rne(H.Q_HOME, z, z, 0.0), oracle.rne(H.Q_HOME, z, z, 0.0), atol=1e-12)
    lim = f["torque_limit"][:6, None]
    np.savez(os.path.join(OUT, "model_override.npz"), seed=seed, model=model, q=q, qd=qd, qdd=qdd, mass=mass,
             tau_rne=tau, feasible_rne=(np.abs(tau[:6]) < lim).all(axis=0).astype(np.uint8),
             tau_nov=tau_static, feasible_nov=(np.abs(tau_static[:6]) < lim).all(axis=0).astype(np.uint8))
    t, ok = oracle.torque_test_batch("rne", q, qd, qdd, mass, model=model)
    print("model_override: oracle vs reference %.2e N.m, feasible %.3f" % (np.abs(t - tau).max(), ok.mean()))


def gen_minjerk(seed=11):
    rng = np.random.default_rng(seed)
    out = {}
    for name, L, n_int in [("p2", 2, 64), ("p5", 5, 37), ("p20", 20, 250), ("p3_1", 3, 1)]:
        pts = rng.uniform(Q_LO, Q_HI, size=(L, 7))
        if L == 5:  # force some via-velocity sign changes and near-zero products (min_jerk_v2.py:118)
            pts[2, 0] = pts[1, 0]
            pts[3, 1] = pts[2, 1] + 1e-6
            pts[2, 1] = pts[1, 1] + 1e-6
        coeffs, x, v, a = H.ref_minjerk(pts, n_int)
        out[name + "_points"] = pts
        out[name + "_n"] = n_int
        out[name + "_coeffs"] = np.transpose(coeffs[:, :, :6], (1, 0, 2))  # -> [seg][k][6]
        out[name + "_x"], out[name + "_v"], out[name + "_a"] = x, v, a
    np.savez(os.path.join(OUT, "minjerk.npz"), **out)
    print("minjerk ok")


def gen_edges(n=160, W=64, seed=4, mass=5.0):
    qa, qb = sample_edges(n, seed)
    ff = np.zeros(n, dtype=np.int32)
    for e in range(n):
        _, x, v, a = H.ref_minjerk(np.stack([qa[:, e], qb[:, e]]), W)
        f = W
        for w in range(W):
            ok, _ = H.ref_torque_test("rne", x[w], v[w], a[w], mass)
            if not ok:
                f = w
                break
        ff[e] = f
    np.savez(os.path.join(OUT, "edges_cfg4.npz"), seed=seed, qa=qa, qb=qb, W=W, mass=mass, first_fail=ff)
    print("edges_cfg4: feasible frac %.3f" % (ff == W).mean())


def gen_traj(seed=5, L=6, T=0.4, mass=5.0, sigma=1.3):
    """rrt_star.py:203-210 final check on a short path: n = int(T*1000/L) samples per segment."""
    rng = np.random.default_rng(seed)
    pts = np.empty((L, 7))
    pts[0] = H.Q_HOME
    for i in range(1, L):
        pts[i] = np.clip(pts[i - 1] + rng.normal(0, sigma, 7), Q_LO, Q_HI)
    n_int = int(T * 1000 / L)
    _, x, v, a = H.ref_minjerk(pts, n_int)
    ns = x.shape[0]
    tau = np.zeros((ns, 7))
    ok = np.zeros(ns, dtype=np.uint8)
    tau_nopayload = np.zeros((ns, 7))  # Conf.__init__ logging pass (utils.py:3376-3378): rne without payload
    for i in range(ns):
        o, t = H.ref_torque_test("rne", x[i], v[i], a[i], mass)
        tau[i], ok[i] = t, o
        tau_nopayload[i] = H.ref_rne(x[i], v[i], a[i], 0.0)
    np.savez(os.path.join(OUT, "traj.npz"), points=pts, n_int=n_int, mass=mass, x=x, v=v, a=a, tau=tau,
             feasible=ok, tau_nopayload=tau_nopayload)
    print("traj: %d samples, feasible %.3f" % (ns, ok.mean()))


def gen_ik(n=1500, n_free=4, seed=3):
    """Config-3 distribution: pose = reference ComputeFk(q), free sweep = own j7 then uniform."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(Q_LO[:, None], Q_HI[:, None], size=(7, n))
    trans, rot = oracle.ref_fk_batch(q)
    free = np.empty((n_free, n))
    free[0] = q[6]
    free[1:] = rng.uniform(-2.8973, 2.8973, size=(n_free - 1, n))
    sols, counts = oracle.ref_ik_batch(rot, trans, free)
    # special poses: q_home and a few axis-aligned / degenerate-looking ones
    qs = np.array([H.Q_HOME, [0.5, 0.9, -0.3, -1.2, 0.7, 2.5, -1.0], [0, 0, 0, -1.5, 0, 1.5, 0],
                   [0, 0, 0, 0, 0, 0, 0], [0.3, 0.0, 0.0, -2.0, 0.0, 2.0, 0.3],
                   [1.0, 0.5, 0.0, -2.0, math.pi / 2, 1.0, 0.5]]).T
    st, sr = oracle.ref_fk_batch(qs)
    sfree = qs[6:7].copy()
    ssols, scounts = oracle.ref_ik_batch(sr, st, sfree)
    np.savez_compressed(os.path.join(OUT, "ik_cfg3.npz"), seed=seed, q=q, trans=trans, rot=rot, free=free, sols=sols,
             counts=counts, special_q=qs, special_trans=st, special_rot=sr, special_free=sfree,
             special_sols=ssols, special_counts=scounts)
    vals, cnt = np.unique(counts, return_counts=True)
    print("ik_cfg3 counts:", dict(zip(vals.tolist(), cnt.tolist())), "special:", scounts.tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    oracle.build()
    which = sys.argv[1:] or ["kats", "states", "model", "minjerk", "edges", "traj", "ik"]
    for w in which:
        {"kats": gen_kats, "states": gen_states, "model": gen_model, "minjerk": gen_minjerk, "edges": gen_edges,
         "traj": gen_traj, "ik": gen_ik}[w]()
