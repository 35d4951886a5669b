"""B200-native batched manipulator kinematics: a drop-in for the data-parallel hot path
of Sarrasor/RoboticManipulators (robots/, utils/ keep the reference's call signatures
and gain *_batch entry points backed by hand-written sm_100a CUDA kernels behind the
C ABI in include/rk_b200.h).  There is no CPU fallback."""
from . import _lib  # noqa: F401

__all__ = ["robots", "utils"]
__version__ = "0.1.0"
