"""CPU: the product's RNE recursion (csrc/panda_model.cuh -- the same source the CUDA kernels instantiate: the
customised 3-vector Newton-Euler, the compile-time base-parameter regrouping, the joint sincos and the run-time
model folding of tcmp_rne_batch_model) built for the host by tests/native/rne_host.cpp and compared with the
oracle and with the golden vectors the reference produced.  A test harness only: libtcmp.so has no CPU path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, load_golden, sample_states

NATIVE = os.path.join(ROOT, "tests", "native")
_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def host_rne():
    so = os.path.join(NATIVE, "librne_host.so")
    src = os.path.join(NATIVE, "rne_host.cpp")
    core = os.path.join(ROOT, "torque_constrained_motion_planning_b200", "csrc", "panda_model.cuh")
    hdr = os.path.join(ROOT, "include", "tcmp.h")
    tab = os.path.join(ROOT, "torque_constrained_motion_planning_b200", "csrc", "sincos_table.inc")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in (src, core, hdr, tab)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, src])
    L = ctypes.CDLL(so)

    def fn(mode, q, qd=None, qdd=None, mass=0.0, threshold=0.01, model=None):
        arr = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        q, qd, qdd = arr(q), arr(qd), arr(qdd)
        n = q.shape[1]
        pm = None if np.ndim(mass) == 0 else arr(mass)
        tau = np.empty((7, n))
        ok = np.empty(n, np.uint8)
        p = lambda a: None if a is None else a.ctypes.data_as(_dp)
        mdl = arr(model)
        L.host_rne_batch(p(mdl), oracle.MODES[mode], ctypes.c_int64(n), p(q), p(qd), p(qdd), p(pm),
                         ctypes.c_double(0.0 if pm is not None else float(mass)), ctypes.c_double(threshold),
                         p(tau), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return tau, ok
    return fn


@pytest.mark.parametrize("mode", ["rne", "nov", "dyn"])
def test_compiled_in_panda_matches_oracle(host_rne, mode):
    q, qd, qdd, mass = sample_states(20000, seed=31)
    tau, ok = host_rne(mode, q, qd, qdd, mass)
    tau_o, ok_o = oracle.torque_test_batch(mode, q, qd, qdd, mass)
    assert np.abs(tau - tau_o).max() < 1e-11
    assert np.array_equal(ok, ok_o)


def test_compiled_in_panda_matches_reference_golden(host_rne):
    g = load_golden("states_cfg2.npz")
    tau, ok = host_rne("rne", g["q"], g["qd"], g["qdd"], g["mass"])
    assert np.abs(tau - g["tau_rne"]).max() < 1e-11 and np.array_equal(ok, g["feasible_rne"])
    tau, ok = host_rne("nov", g["q"], None, None, g["mass"])
    assert np.abs(tau - g["tau_nov"]).max() < 1e-11 and np.array_equal(ok, g["feasible_nov"])


def test_static_call_without_velocities(host_rne):
    q, _, _, mass = sample_states(2000, seed=32)
    tau, ok = host_rne("rne", q, None, None, mass)
    tau_o, ok_o = oracle.torque_test_batch("rne", q, None, None, mass)
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)


def test_large_and_nonfinite_angles_follow_libm(host_rne):
    """Angles beyond 1e5 rad (or non-finite) leave the joint fast sincos for libm's: still the oracle's result."""
    q, qd, qdd, mass = sample_states(512, seed=33)
    q[3, ::7] += 2e5
    q[5, ::11] = 1e300
    tau, ok = host_rne("rne", q, qd, qdd, mass)
    tau_o, ok_o = oracle.torque_test_batch("rne", q, qd, qdd, mass)
    assert np.abs(tau - tau_o).max() < 1e-9 and np.array_equal(ok, ok_o)
    q[2, 0] = np.nan
    tau, ok = host_rne("rne", q, qd, qdd, mass)
    assert np.isnan(tau[:, 0]).any() and ok[0] == 1   # NaN never compares >= limit (panda_primitives.py:182-183)


def test_default_record_is_the_compiled_in_panda(host_rne):
    """Folding + regrouping the default record at run time reproduces the compile-time constants' torques."""
    q, qd, qdd, mass = sample_states(5000, seed=34)
    for mode in ("rne", "nov", "dyn"):
        a, oa = host_rne(mode, q, qd, qdd, mass)
        b, ob = host_rne(mode, q, qd, qdd, mass, model=oracle.default_model())
        assert np.abs(a - b).max() < 1e-12 and np.array_equal(oa, ob)


def test_model_override_matches_reference_golden(host_rne):
    """rne.py executed with its inertial lists overwritten (oracle/make_golden.py gen_model): a massive link8 with an
    off-axis COM, a heavier hand, perturbed links, another payload lever, tighter limits."""
    g = load_golden("model_override.npz")
    tau, ok = host_rne("rne", g["q"], g["qd"], g["qdd"], g["mass"], model=g["model"])
    assert np.abs(tau - g["tau_rne"]).max() < 1e-11 and np.array_equal(ok, g["feasible_rne"])
    tau, ok = host_rne("nov", g["q"], None, None, g["mass"], model=g["model"])
    assert np.abs(tau - g["tau_nov"]).max() < 1e-11 and np.array_equal(ok, g["feasible_nov"])


@pytest.mark.parametrize("mode", ["rne", "nov", "dyn"])
def test_model_override_matches_oracle(host_rne, mode):
    g = load_golden("model_override.npz")
    q, qd, qdd, mass = sample_states(10000, seed=35)
    tau, ok = host_rne(mode, q, qd, qdd, mass, model=g["model"])
    tau_o, ok_o = oracle.torque_test_batch(mode, q, qd, qdd, mass, model=g["model"])
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)


def _table_lib():
    so = os.path.join(NATIVE, "librne_host.so")
    return ctypes.CDLL(so)


def test_table_sincos_accuracy(host_rne):
    """K1's table-driven sincos (csrc/sincos_table.inc + two short polynomials) against extended-precision libm:
    abs error <= 2.5e-16 over the joint range, near multiples of pi/2 (where a relative error would show) and up to
    the 4096 rad switch-over; beyond it (and for inf / nan) the fast path declines."""
    L = _table_lib()
    rng = np.random.default_rng(36)
    half_pi = np.arange(-40, 41) * (np.pi / 2)
    x = np.concatenate([rng.uniform(-3.8, 3.8, 400_000), rng.uniform(-4095.9, 4095.9, 200_000),
                        half_pi, half_pi + 1e-9, half_pi - 3e-13, np.arange(-1024, 1025) * (np.pi / 512),
                        [0.0, -0.0, 5e-324, 1e-300, 4095.999]])
    s, c = np.empty_like(x), np.empty_like(x)
    ok = np.empty(len(x), np.uint8)
    p = lambda a: a.ctypes.data_as(_dp)
    L.host_sincos_table(ctypes.c_int64(len(x)), p(x), p(s), p(c), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert ok.all()
    xl = x.astype(np.longdouble)
    es = np.abs(s - np.sin(xl)).max()
    ec = np.abs(c - np.cos(xl)).max()
    assert es < 2.5e-16 and ec < 2.5e-16, (es, ec)
    assert np.abs(s * s + c * c - 1).max() < 5e-16
    far = np.array([4096.0, -5000.0, 1e300, np.inf, -np.inf, np.nan])
    s2, c2, ok2 = np.empty_like(far), np.empty_like(far), np.empty(len(far), np.uint8)
    L.host_sincos_table(ctypes.c_int64(len(far)), p(far), p(s2), p(c2), ok2.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert not ok2.any()


def test_table_path_torques_match_oracle_and_reference(host_rne):
    L = _table_lib()
    q, qd, qdd, mass = sample_states(50_000, seed=37)
    q[4, ::97] += 5000.0                              # some states take the libm fallback
    tau = np.empty((7, q.shape[1]))
    ok = np.empty(q.shape[1], np.uint8)
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(_dp)
    L.host_rne_batch_table(ctypes.c_int64(q.shape[1]), p(q), p(qd), p(qdd), p(mass), ctypes.c_double(0.01), p(tau),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    tau_o, ok_o = oracle.torque_test_batch("rne", q, qd, qdd, mass)
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)
    g = load_golden("states_cfg2.npz")
    n = g["q"].shape[1]
    tau, ok = np.empty((7, n)), np.empty(n, np.uint8)
    L.host_rne_batch_table(ctypes.c_int64(n), p(g["q"]), p(g["qd"]), p(g["qdd"]), p(g["mass"]), ctypes.c_double(0.01),
                           p(tau), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert np.abs(tau - g["tau_rne"]).max() < 1e-11 and np.array_equal(ok, g["feasible_rne"])


def test_structural_properties_of_the_recursion(host_rne):
    """Size-independent properties the domain offers, on the product's own recursion (host build):
    joint 1's angle never matters (gravity is along its axis: bit-identical), zero velocities through the dynamic
    path equal the static path, torques are affine in the payload mass, and doubling every mass and inertia of the
    model doubles the static torques."""
    q, qd, qdd, _ = sample_states(4000, seed=38)
    q2 = q.copy()
    q2[0] = np.random.default_rng(39).uniform(-2.8, 2.8, q.shape[1])
    a, _ = host_rne("rne", q, qd, qdd, 3.0)
    b, _ = host_rne("rne", q2, qd, qdd, 3.0)
    assert np.array_equal(a, b)
    z = np.zeros_like(q)
    dyn0, _ = host_rne("rne", q, z, z, 3.0)
    stat, _ = host_rne("rne", q, None, None, 3.0)
    nov, _ = host_rne("nov", q, qd, qdd, 3.0)
    assert np.abs(dyn0 - stat).max() < 1e-12 and np.array_equal(stat, nov)
    t1, _ = host_rne("rne", q, qd, qdd, 1.0)
    t3, _ = host_rne("rne", q, qd, qdd, 3.0)
    t5, _ = host_rne("rne", q, qd, qdd, 5.0)
    assert np.abs(t1 + t5 - 2 * t3).max() < 1e-11
    m = oracle.default_model()
    f = oracle.model_fields(m)
    base, _ = host_rne("nov", q, None, None, 0.0, model=m)
    f["mass"] *= 2.0
    f["inertia"] *= 2.0
    twice, _ = host_rne("nov", q, None, None, 0.0, model=m)
    assert np.abs(twice - 2 * base).max() < 1e-11
    # dyn mode: the payload enters only through the tool-point force, so tau(dyn, m) - tau(rne, 0) is linear in m
    d0, _ = host_rne("dyn", q, qd, qdd, 0.0)
    d2, _ = host_rne("dyn", q, qd, qdd, 2.0)
    d4, _ = host_rne("dyn", q, qd, qdd, 4.0)
    r0, _ = host_rne("rne", q, qd, qdd, 0.0)
    assert np.abs(d0 - r0).max() < 1e-12 and np.abs((d4 - d0) - 2 * (d2 - d0)).max() < 1e-11


def test_sincos_table_is_what_the_generator_produces():
    """csrc/sincos_table.inc is generated (scripts/make_sincos_table.py, 70-digit arithmetic): regenerate and compare
    every entry bit for bit; exact zeros / ones sit at the quadrant points."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_sincos_table", os.path.join(ROOT, "scripts", "make_sincos_table.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    path = os.path.join(ROOT, "torque_constrained_motion_planning_b200", "csrc", "sincos_table.inc")
    rows = [l.strip().strip("{},").split(", ") for l in open(path) if l.startswith("{")]
    assert len(rows) == 1024
    for i in range(0, 1024, 7):      # every 7th entry keeps the test under a second; the quadrant points below too
        s, c = gen.entry(i)
        assert (float.fromhex(rows[i][0]), float.fromhex(rows[i][1])) == (s, c), i
    for i, (s, c) in {0: (0.0, 1.0), 256: (1.0, 0.0), 512: (0.0, -1.0), 768: (-1.0, 0.0)}.items():
        assert (float.fromhex(rows[i][0]), float.fromhex(rows[i][1])) == (s, c)


def test_special_joint_values(host_rne):
    """Joint values snapped to multiples of pi/4 and to the joint limits (exact table nodes of the table-driven sincos,
    zeros of sin / cos): the polynomial path, the table path and the run-time model path all stay on the oracle."""
    from conftest import Q_HI, Q_LO
    rng = np.random.default_rng(40)
    q, qd, qdd, mass = sample_states(30_000, seed=41)
    snap = rng.random(q.shape) < 0.5
    q = np.where(snap, np.clip(np.round(q / (np.pi / 4)) * (np.pi / 4), Q_LO[:, None], Q_HI[:, None]), q)
    tau_o, ok_o = oracle.torque_test_batch("rne", q, qd, qdd, mass)
    tau, ok = host_rne("rne", q, qd, qdd, mass)
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)
    tau, ok = host_rne("rne", q, qd, qdd, mass, model=oracle.default_model())
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)
    L = _table_lib()
    n = q.shape[1]
    tau, ok = np.empty((7, n)), np.empty(n, np.uint8)
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(_dp)
    L.host_rne_batch_table(ctypes.c_int64(n), p(q), p(qd), p(qdd), p(mass), ctypes.c_double(0.01), p(tau),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert np.abs(tau - tau_o).max() < 1e-11 and np.array_equal(ok, ok_o)
