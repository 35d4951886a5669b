"""ctypes binding of librk_b200.so (the C ABI declared in include/rk_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, the
caller gets an exception.  torch is used only for device memory and streams.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librk_b200.so")

RK_AOS, RK_SOA, RK_JAC_USER_ANGLES = 0, 1, 2
RK_MAX_TOKENS = 32
RK_CALIB_MAX_JOINTS, RK_CALIB_MAX_PARAMS, RK_CALIB_MAX_TOOLS = 8, 19, 4

c_double_p = C.POINTER(C.c_double)


class RkError(RuntimeError):
    pass


class RkFanuc(C.Structure):
    _fields_ = [("l", C.c_double * 7), ("offsets", C.c_double * 6),
                ("directions", C.c_double * 6), ("lim_lo", C.c_double * 6),
                ("lim_hi", C.c_double * 6), ("d", C.c_double), ("dq", C.c_double),
                ("eps", C.c_double), ("T_base", C.c_double * 12),
                ("T_tool", C.c_double * 12), ("T_base_inv", C.c_double * 12),
                ("T_tool_inv", C.c_double * 12), ("has_base", C.c_int32),
                ("has_tool", C.c_int32)]


class RkArm(C.Structure):
    _fields_ = [("l", C.c_double * 3), ("lim_lo", C.c_double * 3),
                ("lim_hi", C.c_double * 3), ("eps", C.c_double),
                ("T_base", C.c_double * 12), ("T_tool", C.c_double * 12),
                ("T_base_inv", C.c_double * 12), ("T_tool_inv", C.c_double * 12),
                ("has_base", C.c_int32), ("has_tool", C.c_int32)]


class RkChain(C.Structure):
    _fields_ = [("ntok", C.c_int32), ("op", C.c_int32 * RK_MAX_TOKENS),
                ("active", C.c_int32 * RK_MAX_TOKENS), ("src", C.c_int32 * RK_MAX_TOKENS),
                ("cst", C.c_double * RK_MAX_TOKENS), ("T_base", C.c_double * 12),
                ("T_tool", C.c_double * 12), ("has_base", C.c_int32),
                ("has_tool", C.c_int32)]


class RkCalibModel(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("nparams", C.c_int32), ("ntools", C.c_int32),
                ("joint_axis", C.c_int32 * RK_CALIB_MAX_JOINTS),
                ("joint_kind", C.c_int32 * RK_CALIB_MAX_JOINTS),
                ("joint_dir", C.c_double * RK_CALIB_MAX_JOINTS),
                ("joint_off", C.c_double * RK_CALIB_MAX_JOINTS),
                ("seg", (C.c_double * 12) * RK_CALIB_MAX_JOINTS),
                ("param_joint", C.c_int32 * RK_CALIB_MAX_PARAMS),
                ("param_kind", C.c_int32 * RK_CALIB_MAX_PARAMS),
                ("param_axis", (C.c_double * 3) * RK_CALIB_MAX_PARAMS),
                ("param_origin", (C.c_double * 3) * RK_CALIB_MAX_PARAMS),
                ("T_base", C.c_double * 12),
                ("tool_p", (C.c_double * 3) * RK_CALIB_MAX_TOOLS)]


_vp, _i64, _int = C.c_void_p, C.c_int64, C.c_int
_SIGNATURES = {
    "rk_version": (C.c_int, []),
    "rk_strerror": (C.c_char_p, [_int]),
    "rk_last_cuda_error": (C.c_char_p, []),
    "rk_fanuc_fk_jac_f64": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "rk_fanuc_fk_jac_f32": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "rk_fanuc_fk_jac_ik8_f64": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_fanuc_fk_jac_ik8_f32": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_fanuc_ik_f64": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_fanuc_ik_f32": (_int, [C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_host_pipe_create": (_int, [C.POINTER(C.c_void_p), _i64, _int]),
    "rk_host_pipe_destroy": (_int, [_vp]),
    "rk_host_pipe_last_ms": (C.c_float, [_vp]),
    "rk_fanuc_init": (_int, [C.POINTER(RkFanuc), C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rk_arm_init": (_int, [C.POINTER(RkArm), _int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_double)]),
    "rk_fanuc_fk_jac_ik8_host_f64": (_int, [_vp, C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    "rk_fanuc_fk_jac_ik8_host_f32": (_int, [_vp, C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    "rk_fanuc_ik8_host_f64": (_int, [_vp, C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp]),
    "rk_fanuc_ik8_host_f32": (_int, [_vp, C.POINTER(RkFanuc), _vp, _i64, _i64, _int, _vp, _vp]),
    "rk_rrr_fk_jac_f64": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_rrr_ik_f64": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_rr_fk_jac_f64": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_rr_ik_f64": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_rrr_fk_jac_f32": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_rrr_ik_f32": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_rr_fk_jac_f32": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_rr_ik_f32": (_int, [C.POINTER(RkArm), _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "rk_rrr_p2p_fk_f32": (_int, [C.POINTER(RkArm), _vp, _vp, _i64, _int, C.c_double, C.c_double,
                                 C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_chain_fk_jac_f64": (_int, [C.POINTER(RkChain), _vp, _int, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_chain_fk_jac_f32": (_int, [C.POINTER(RkChain), _vp, _int, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_chain_jit_status": (C.c_char_p, []),
    "rk_chain_jit_source": (_i64, [C.POINTER(RkChain), _int, _int, _int, _int, C.c_char_p, _i64]),
    "rk_chain_dq_from_dx_f64": (_int, [C.POINTER(RkChain), _vp, _int, C.POINTER(C.c_int32), _vp, _i64, _i64,
                                       _int, _vp, _vp, _vp]),
    "rk_fanuc_dq_from_dx_f64": (_int, [C.POINTER(RkFanuc), _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_fanuc_dq_from_dx_f32": (_int, [C.POINTER(RkFanuc), _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "rk_trap_coefficients_f64": (_int, [_vp, _vp, _i64, C.c_double, C.c_double, C.c_double,
                                        _vp, _vp, _vp]),
    "rk_trap_sample_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp,
                                  _vp]),
    "rk_poly5_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    "rk_trap_p2p_f64": (_int, [_vp, _vp, _i64, _int, _int, C.c_double, C.c_double, C.c_double,
                               _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_rrr_p2p_fk_f64": (_int, [C.POINTER(RkArm), _vp, _vp, _i64, _int, C.c_double, C.c_double,
                                 C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_rrr_lin_ik_f64": (_int, [C.POINTER(RkArm), _vp, _vp, _i64, _int, C.c_double, C.c_double,
                                 C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rk_calib_out_len": (_int, [_int]),
    "rk_calib_workspace_bytes": (C.c_size_t, []),
    "rk_calib_normal_eq_f64": (_int, [C.POINTER(RkCalibModel), _vp, _vp, _i64, _vp, _vp,
                                      C.c_size_t, _vp]),
    "rk_calib_eval_f64": (_int, [C.POINTER(RkCalibModel), _vp, _vp, _i64, _vp, _vp,
                                 C.c_size_t, _vp]),
    "rk_calib_tool_base_eq_f64": (_int, [C.POINTER(RkCalibModel), _vp, _vp, _i64, _vp, _vp,
                                         C.c_size_t, _vp]),
    "rk_probe_dfma": (_int, [_i64, _vp, C.POINTER(C.c_int64), _vp]),
    "rk_probe_dmma": (_int, [_i64, _vp, C.POINTER(C.c_int64), _vp]),
    "rk_probe_atan2": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "rk_probe_stream": (_int, [_vp, _int, _vp, _int, _i64, _i64, _int, _int, _int, _vp]),
    "rk_probe_sincos": (_int, [_vp, _i64, _vp, _vp, _vp]),
}

_lib = None


def exported_symbols():
    """Names every build of the library must export (checked by the CPU tests)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen librk_b200.so and set the prototypes; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RkError(
            f"{LIB_PATH} is missing: build it with "
            "`python roboticmanipulators_b200/csrc/build.py` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    missing = [name for name in _SIGNATURES if not hasattr(lib, name)]
    if missing:
        raise RkError(f"{LIB_PATH} is stale: missing symbols {missing}; rebuild it")
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        lib = load()
        msg = lib.rk_strerror(status).decode()
        if status == -2:
            msg += ": " + lib.rk_last_cuda_error().decode()
        raise RkError(f"librk_b200: {msg} (status {status})")


def ptr(t):
    """Device pointer of a torch tensor (or None) -- a plain int: ctypes converts it for a
    c_void_p parameter, and skipping the c_void_p object saves ~0.4 us per argument."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name="tensor"):
    if not getattr(t, "is_cuda", False):
        raise RkError(f"{name} must be a torch CUDA tensor (no CPU fallback)")


def xf12(T):
    """4x4 array-like -> 12 doubles (3x4 row-major)."""
    A = np.array(T, dtype=np.float64).reshape(4, 4)
    return A[:3, :].reshape(12)


def fill(carray, values):
    vals = np.asarray(values, dtype=np.float64).reshape(-1)
    for i, v in enumerate(vals):
        carray[i] = float(v)


def is_identity(T):
    A = np.array(T, dtype=np.float64).reshape(4, 4)
    return bool(np.array_equal(A, np.eye(4)))
