"""CPU: libtcmp.so loads and exports every symbol include/tcmp.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np

import oracle
import pytest

from conftest import ROOT

from torque_constrained_motion_planning_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tcmp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcmp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    from torque_constrained_motion_planning_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.tcmp_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libtcmp.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call, so these run on a CPU-only box."""
    lib = _lib.load()
    assert lib.tcmp_rne_batch(9, 0, 1, None, None, None, None, 0.0, 0.01, None, None, None) == -1
    assert b"bad mode" in lib.tcmp_last_error()
    assert lib.tcmp_rne_batch(0, 7, 1, None, None, None, None, 0.0, 0.01, None, None, None) == -1
    assert lib.tcmp_rne_batch(0, 0, -5, None, None, None, None, 0.0, 0.01, None, None, None) == -1
    assert lib.tcmp_rne_batch(0, 0, 4, None, None, None, None, 0.0, 0.01, None, None, None) == -1  # q NULL
    assert lib.tcmp_rne_batch(0, 0, 0, None, None, None, None, 0.0, 0.01, None, None, None) == 0   # empty batch
    assert lib.tcmp_edge_feasibility(0, 0, 4, 0, None, None, 0.0, 0.01, 0, None, None) == -1
    assert lib.tcmp_ik_batch(4, None, None, None, 1, 0, None, None, None, None) == -1
    assert lib.tcmp_ik_batch(0, None, None, None, 1, 0, None, None, None, None) == 0


def test_model_record_without_gpu():
    """tcmp_model_default fills the reference's tables (no CUDA call), and a malformed record is rejected before any."""
    from torque_constrained_motion_planning_b200 import engine
    m = engine.InertialModel.default()
    assert np.array_equal(m.record, oracle.default_model())
    assert m.mass[8] == 0.68 and m.payload_radius == 0.14 + 0.025 and m.tool_z == 0.105
    lib = _lib.load()
    bad = engine.InertialModel.default()
    bad.mass[3] = -1.0
    assert lib.tcmp_rne_batch_model(bad.record.ctypes.data, 0, 0, 0, None, None, None, None, 0.0, 0.01, None, None,
                                    None) == -1
    assert b"mass" in lib.tcmp_last_error()
    bad = engine.InertialModel.default()
    bad.torque_limit[2] = 0.0
    assert lib.tcmp_rne_batch_model(bad.record.ctypes.data, 0, 0, 0, None, None, None, None, 0.0, 0.01, None, None,
                                    None) == -1
    bad.torque_limit[2] = float("nan")
    assert lib.tcmp_rne_batch_model(bad.record.ctypes.data, 0, 0, 0, None, None, None, None, 0.0, 0.01, None, None,
                                    None) == -1
    ok = engine.InertialModel.default()
    assert lib.tcmp_rne_batch_model(ok.record.ctypes.data, 0, 0, 0, None, None, None, None, 0.0, 0.01, None, None,
                                    None) == 0
    assert lib.tcmp_model_default(None) == -1


def test_limits_table():
    from torque_constrained_motion_planning_b200 import engine
    lim = engine.get_limits()
    assert lim["torque"].tolist() == [87, 87, 87, 87, 12, 12, 12]
    assert np.all(lim["q_lo"] < lim["q_hi"])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.TcmpError, match="no CPU fallback"):
        _lib.load()


def test_header_is_plain_c(tmp_path):
    """include/tcmp.h is a C interface: it must compile as C99 (and as C++11) on its own, warnings as errors, and a C
    program linked against libtcmp.so must resolve and run the CUDA-free entry points."""
    import subprocess
    from conftest import ROOT
    src = tmp_path / "use_tcmp.c"
    src.write_text('#include <stdio.h>\n#include "tcmp.h"\n'
                   'int main(void) {\n'
                   '    tcmp_model m; double lim[7];\n'
                   '    if (tcmp_abi_version() != TCMP_ABI_VERSION) return 1;\n'
                   '    if (tcmp_model_default(&m) != 0 || m.mass[8] != 0.68) return 2;\n'
                   '    if (tcmp_get_limits(lim, 0, 0, 0) != 0 || lim[0] != 87.0) return 3;\n'
                   '    if (tcmp_rne_batch(9, 0, 1, 0, 0, 0, 0, 0.0, 0.01, 0, 0, 0) == 0) return 4;\n'
                   '    printf("%s\\n", tcmp_last_error());\n'
                   '    return 0;\n}\n')
    inc = os.path.join(ROOT, "include")
    for cc, std in (("/usr/bin/gcc", "-std=c99"), ("/usr/bin/g++", "-std=c++11")):
        subprocess.check_call([cc, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc] +
                              (["-x", "c++"] if cc.endswith("g++") else []) + [str(src)])
    exe = tmp_path / "use_tcmp"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ltcmp",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "bad mode" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_product_never_reaches_into_the_oracle():
    """The checker is test infrastructure: no module of the product imports it, no native source includes it, and
    importing the whole package leaves `oracle` out of sys.modules (run in a fresh interpreter)."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "torque_constrained_motion_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
            assert not re.search(r'#include\s+"[^"]*oracle', text), f
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import torque_constrained_motion_planning_b200 as p\n"
            "from torque_constrained_motion_planning_b200 import (engine, rne, ikfast_panda_arm, ik_utils, ikfast,\n"
            "    franka_ik_fast, min_jerk_v2, utils, collision, rrt_star, panda_primitives, panda_model, distributed)\n"
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'\n" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_library_carries_sm_100a_code_for_every_kernel_family():
    """libtcmp.so embeds native sm_100a cubins (no PTX-only JIT path, no other architectures) with the kernels the
    design names."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    cubins = re.findall(r"(\w+)\.(sm_\w+)\.cubin", elfs)
    assert cubins and {arch for _, arch in cubins} == {"sm_100a"}, elfs
    usage = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for kernel in ("rne_batch_kernel", "rne_model_kernel", "edge_kernel", "traj_kernel", "ik_kernel_cta", "ik_redo_kernel",
                   "fk_kernel", "ik_select_kernel", "collision_kernel", "extend_prefix_kernel", "peer_signal_kernel",
                   "peer_wait_kernel", "peer_push_kernel", "fp64_peak_kernel"):
        assert kernel in usage, kernel
