"""Robot base class -- mirrors robots/robot.py:9-63 of the reference."""
from abc import ABC, abstractmethod

import numpy as np
import sympy as sp

from .. import _lib


class Robot(ABC):
    """Abstract robot.  `epsilon` is the float-equality threshold (robot.py:17)."""

    epsilon = 1e-5

    @abstractmethod
    def forward_kinematics(self, q_values):
        """Pose T from joint values."""

    @abstractmethod
    def inverse_kinematics(self, T):
        """Joint values from pose T."""

    def set_transforms(self, T_base=None, T_tool=None):
        """World->base and flange->tool transforms; identity when None
        (robot.py:39-57).  Kept as sympy matrices like the reference."""
        self.T_base = sp.eye(4) if T_base is None else sp.Matrix(T_base)
        self.T_tool = sp.eye(4) if T_tool is None else sp.Matrix(T_tool)

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
