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
