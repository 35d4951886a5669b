"""Robot base class -- mirrors robots/robot.py:9-63 of the reference."""
from abc import ABC, abstractmethod

import numpy as np
import sympy as sp
import torch

from .. import _lib


class Robot(ABC):
    """Abstract robot.  `epsilon` is the float-equality threshold (robot.py:17).

    The batched entry points hand the kernels a small constant block (RkFanuc / RkArm, include/
    rk_b200.h).  Building it means converting SymPy matrices to NumPy and inverting them (~80 us),
    so it is built once and kept: assigning ANY attribute of the robot drops it, and a cheap
    fingerprint of the array-valued attributes catches in-place edits (`robot.qs_offsets[0] = ..`).
    T_base / T_tool are immutable SymPy matrices for the same reason: replace them (set_transforms
    or plain assignment), do not edit their entries."""

    epsilon = 1e-5

    @abstractmethod
    def forward_kinematics(self, q_values):
        """Pose T from joint values."""

    @abstractmethod
    def inverse_kinematics(self, T):
        """Joint values from pose T."""

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name != "_c_struct":
            object.__setattr__(self, "_c_struct", None)

    def set_transforms(self, T_base=None, T_tool=None):
        """World->base and flange->tool transforms; identity when None
        (robot.py:39-57).  Kept as sympy matrices like the reference."""
        self.T_base = sp.ImmutableMatrix(sp.eye(4) if T_base is None else sp.Matrix(T_base))
        self.T_tool = sp.ImmutableMatrix(sp.eye(4) if T_tool is None else sp.Matrix(T_tool))

    def set_lengths(self, lengths):
        self._ls = (0.8, 0.8) if lengths is None else lengths   # robot.py:59-63

    # -- helpers shared by the batched entry points ---------------------------------
    def _fill_transforms(self, st):
        """Write T_base/T_tool and their inverses into a ctypes parameter struct."""
        Tb = np.array(self.T_base, dtype=np.float64)
        Tt = np.array(self.T_tool, dtype=np.float64)
        _lib.fill(st.T_base, _lib.xf12(Tb))
        _lib.fill(st.T_tool, _lib.xf12(Tt))
        _lib.fill(st.T_base_inv, _lib.xf12(np.linalg.inv(Tb)))
        _lib.fill(st.T_tool_inv, _lib.xf12(np.linalg.inv(Tt)))
        st.has_base = 0 if _lib.is_identity(Tb) else 1
        st.has_tool = 0 if _lib.is_identity(Tt) else 1

    def _fingerprint(self):
        """Bytes of the array-valued constants a caller may edit in place."""
        return b""

    def _struct(self):
        """Build the C-ABI constant block from the current attributes (no caching)."""
        raise NotImplementedError

    def _cached_struct(self):
        """(ctypes.byref of) the constant block, rebuilt only when an attribute changed."""
        c = self.__dict__.get("_c_struct")
        key = (self.epsilon, self._fingerprint())
        if c is None or c[0] != key or c[1] is not self.T_base or c[2] is not self.T_tool:
            st = self._struct()
            import ctypes as C
            c = (key, self.T_base, self.T_tool, st, C.byref(st))
            object.__setattr__(self, "_c_struct", c)
        return c[4]


try:                                      # raw handle of the current stream without building a Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                    # pragma: no cover
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def launch(fn, dev, *args):
    """Call a C-ABI entry point with `dev` current (kernels launch on the current device) and the
    current torch stream of that device appended as the last argument."""
    cur = torch.cuda.current_device()
    if dev.index is None or dev.index == cur:
        rc = fn(*args, _raw_stream(cur))
        if rc:
            _lib.check(rc)
    else:
        with torch.cuda.device(dev):
            _lib.check(fn(*args, _raw_stream(dev.index)))
