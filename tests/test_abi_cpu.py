"""CPU checks of the drop-in boundary: the library loads, exports every symbol declared in
include/rk_b200.h, the ctypes mirrors agree with the C structs, and argument validation
works without touching a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rk_b200.h")


def _lib():
    import __graft_entry__ as g
    g.build()
    from roboticmanipulators_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol():
    L = _lib()
    lib = L.load()
    with open(HEADER) as f:
        text = f.read()
    declared = set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations found"
    assert declared == set(L.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rk_version() == 100
    assert lib.rk_strerror(0) == b"ok"
    assert lib.rk_strerror(-1) == b"invalid argument"
    assert lib.rk_calib_out_len(18) == 190
    assert lib.rk_calib_workspace_bytes() > 0


def test_ctypes_structs_match_header(tmp_path):
    L = _lib()
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rk_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(RkFanuc), '
                   'sizeof(RkArm), sizeof(RkChain), sizeof(RkCalibModel), '
                   'offsetof(RkFanuc, has_base), offsetof(RkChain, cst), '
                   'offsetof(RkCalibModel, param_axis), offsetof(RkCalibModel, tool_p));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(L.RkFanuc), C.sizeof(L.RkArm), C.sizeof(L.RkChain), C.sizeof(L.RkCalibModel),
            L.RkFanuc.has_base.offset, L.RkChain.cst.offset, L.RkCalibModel.param_axis.offset,
            L.RkCalibModel.tool_p.offset]
    assert got == want


def test_argument_validation_without_gpu():
    L = _lib()
    lib = L.load()
    st = L.RkFanuc()
    # negative sizes / missing branch list are rejected before any CUDA call
    assert lib.rk_fanuc_fk_jac_f64(C.byref(st), None, -1, 0, 0, None, None, None, None) == -1
    assert lib.rk_fanuc_fk_jac_f64(None, None, 4, 4, 0, None, None, None, None) == -1
    assert lib.rk_fanuc_fk_jac_f64(C.byref(st), None, 0, 0, 0, None, None, None, None) == 0
    assert lib.rk_fanuc_ik_f64(C.byref(st), None, 4, 4, 0, None, 5, None, None, None) == -1
    # the single FK -> IK call: same rules, and T, q_out, valid are required
    args = (None, None, None, None, None, None)
    assert lib.rk_fanuc_fk_jac_ik8_f64(C.byref(st), None, -1, 0, 1, *args) == -1
    assert lib.rk_fanuc_fk_jac_ik8_f64(None, None, 4, 4, 1, *args) == -1
    assert lib.rk_fanuc_fk_jac_ik8_f64(C.byref(st), None, 0, 0, 1, *args) == 0
    assert lib.rk_fanuc_fk_jac_ik8_f32(C.byref(st), C.c_void_p(16), 4, 4, 1, *args) == -1
    assert lib.rk_fanuc_fk_jac_ik8_f64(C.byref(st), C.c_void_p(16), 8, 4, 1, C.c_void_p(16), None, None,
                                       C.c_void_p(16), C.c_void_p(16), None) == -1      # ld < n
    with pytest.raises(L.RkError):
        import torch
        from roboticmanipulators_b200.robots import Fanuc165F
        Fanuc165F().fk_jac_ik_batch(torch.zeros(4, 6, dtype=torch.float64))
    ch = L.RkChain()
    ch.ntok = 40
    assert lib.rk_chain_fk_jac_f64(C.byref(ch), None, 0, 0, 0, 0, None, None, None) == -3
    assert lib.rk_trap_p2p_f64(None, None, 4, 9, 10, 1.0, 1.0, 0.0, None, None, None, None,
                               C.c_void_p(8), None) == -3
    # host-array entry points: no pipe / no robot / negative size are rejected before any CUDA call
    assert lib.rk_fanuc_fk_jac_ik8_host_f64(None, C.byref(st), None, 4, 4, 0, None, None, None, None, None) == -1
    assert lib.rk_fanuc_ik8_host_f64(None, C.byref(st), None, 4, 4, 0, None, None) == -1
    assert lib.rk_fanuc_ik8_host_f32(None, None, None, -1, 0, 0, None, None) == -1
    pipe = C.c_void_p()
    assert lib.rk_host_pipe_create(C.byref(pipe), 16, 3) == -1          # chunk < 32 poses
    assert lib.rk_host_pipe_create(C.byref(pipe), 4096, 1) == -1        # fewer than two slots
    assert lib.rk_host_pipe_create(None, 4096, 3) == -1
    assert lib.rk_host_pipe_destroy(None) == 0
    # dq = J^-1 dx: the row list and the chain are validated on the host
    ch2 = L.RkChain()
    ch2.ntok = 2
    ch2.op[0], ch2.op[1] = 5, 0                                         # Rz, Tx
    ch2.active[0], ch2.active[1] = 1, 1
    ch2.src[0], ch2.src[1] = 0, 1
    rows_bad = (C.c_int32 * 2)(0, 0)                                    # the same row twice
    assert lib.rk_chain_dq_from_dx_f64(C.byref(ch2), None, 2, rows_bad, None, 4, 4, 0, None, None, None) == -1
    rows_oob = (C.c_int32 * 2)(0, 6)
    assert lib.rk_chain_dq_from_dx_f64(C.byref(ch2), None, 2, rows_oob, None, 4, 4, 0, None, None, None) == -1
    assert lib.rk_chain_dq_from_dx_f64(C.byref(ch2), None, 2, None, None, 0, 0, 0, None, None, None) == 0
    assert lib.rk_chain_dq_from_dx_f64(None, None, 2, None, None, 4, 4, 0, None, None, None) == -1


def test_no_cpu_fallback():
    """Batched entry points refuse CPU tensors; the product never imports the oracle."""
    import torch
    L = _lib()
    from roboticmanipulators_b200.robots import Fanuc165F, RRRRobot
    with pytest.raises(L.RkError):
        Fanuc165F().forward_kinematics_batch(torch.zeros(4, 6, dtype=torch.float64))
    with pytest.raises(L.RkError):
        RRRRobot().inverse_kinematics_batch(torch.zeros(4, 4, 4, dtype=torch.float64))
    pkg = os.path.join(ROOT, "roboticmanipulators_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, fn)) as f:
                    assert "oracle" not in f.read(), fn


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    L = _lib()
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.RkError, match="no CPU fallback"):
        L.load()


def test_c_initialiser_matches_the_python_struct():
    """rk_fanuc_init (plain-C callers) fills RkFanuc exactly like Fanuc165F._struct: byte for byte
    for the stock robot, and for custom lengths / offsets / directions with base and tool transforms
    (inverses to rounding: the Python layer uses numpy.linalg.inv)."""
    import ctypes as C
    import numpy as np
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    lib = _lib.load()
    st = _lib.RkFanuc()
    assert lib.rk_fanuc_init(C.byref(st), None, None, None, None, None) == 0
    ref = Fanuc165F()._struct()
    assert bytes(st) == bytes(ref)

    def arr(x):
        a = np.ascontiguousarray(x, dtype=np.float64)
        return a, a.ctypes.data_as(C.POINTER(C.c_double))
    lengths = (300.0, 310.0, 290.0, 1000.0, 200.0, 1300.0, 230.0)
    offsets = np.array([0.1, -1.4, 1.6, 0.2, -0.1, 0.3])
    directions = np.array([1.0, -1.0, -1.0, 1.0, -1.0, 1.0])
    c, s = np.cos(0.3), np.sin(0.3)
    Tb = np.array([[c, -s, 0, 100.0], [s, c, 0, -50.0], [0, 0, 1, 25.0], [0, 0, 0, 1]])
    Tt = np.array([[1, 0, 0, 10.0], [0, c, -s, 0.0], [0, s, c, 120.0], [0, 0, 0, 1]])
    keep = [arr(lengths), arr(offsets), arr(directions), arr(Tb), arr(Tt)]
    assert lib.rk_fanuc_init(C.byref(st), *[k[1] for k in keep]) == 0
    ref = Fanuc165F(T_base=Tb, T_tool=Tt, lengths=lengths, offsets=offsets, directions=directions)._struct()
    for name, _ in _lib.RkFanuc._fields_:
        a, b = getattr(st, name), getattr(ref, name)
        a = np.array(a[:]) if hasattr(a, "__len__") else np.array(a)
        b = np.array(b[:]) if hasattr(b, "__len__") else np.array(b)
        if name.endswith("_inv"):
            assert np.allclose(a, b, rtol=1e-14, atol=1e-12), name
        else:
            assert np.array_equal(a, b), name
    assert lib.rk_fanuc_init(None, None, None, None, None, None) != 0


def test_c_demo_compiles_against_the_header(tmp_path):
    """examples/c_abi_demo.c is plain C99 against include/rk_b200.h and links with the library and
    cudart only (it is run on the GPU box by tests/test_examples_gpu.py)."""
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA headers not available")
    exe = tmp_path / "c_abi_demo"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(root, "examples", "c_abi_demo.c"), "-L", os.path.join(root, "roboticmanipulators_b200"),
                        "-lrk_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_c_arm_initialiser_matches_the_python_structs():
    """rk_arm_init == RRRRobot()._struct() / RRRobot()._struct(), byte for byte."""
    import ctypes as C
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    lib = _lib.load()
    st = _lib.RkArm()
    assert lib.rk_arm_init(C.byref(st), 3, None, None, None) == 0
    assert bytes(st) == bytes(RRRRobot()._struct())
    assert lib.rk_arm_init(C.byref(st), 2, None, None, None) == 0
    assert bytes(st) == bytes(RRRobot()._struct())
    ls = np.array([0.5, 1.25])
    assert lib.rk_arm_init(C.byref(st), 2, ls.ctypes.data_as(C.POINTER(C.c_double)), None, None) == 0
    assert bytes(st) == bytes(RRRobot(lengths=(0.5, 1.25))._struct())
    assert lib.rk_arm_init(C.byref(st), 4, None, None, None) != 0
