"""Shared machinery of the two small serial arms (RRRRobot, RRRobot)."""
import ctypes as C

import numpy as np
import torch

from .. import _batch, _lib
from .robot import Robot, launch


_SFX = {torch.float64: "f64", torch.float32: "f32"}


def _dtype_of(x):
    """float32 inputs run the FP32 kernels; everything else is computed in float64."""
    return x.dtype if x.dtype in _SFX else torch.float64


class _SerialArm(Robot):
    _nq = 3            # number of joints
    _fk_fn = "rk_rrr_fk_jac_"      # + "f64" / "f32"
    _ik_fn = "rk_rrr_ik_"
    _canonical = ((1, 1), (1, -1), (-1, 1), (-1, -1))

    def _lengths(self):
        raise NotImplementedError

    def _calculate_limits_radians(self):
        self.qs_lim_rad = tuple(
            (np.deg2rad(x[0]), np.deg2rad(x[1])) for x in self.qs_lim_deg)

    def _generate_value_pairs(self):
        ls = self._lengths()
        self._value_pairs = [(f"l_{i}", ls[i]) for i in range(len(ls))]

    def _fingerprint(self):
        return np.asarray(self._lengths(), dtype=np.float64).tobytes()

    def _struct(self):
        st = _lib.RkArm()
        ls = list(self._lengths()) + [0.0] * (3 - len(self._lengths()))
        _lib.fill(st.l, ls)
        lim = np.zeros((3, 2))
        lim[:self._nq] = np.asarray(self.qs_lim_rad, dtype=np.float64)
        _lib.fill(st.lim_lo, lim[:, 0])
        _lib.fill(st.lim_hi, lim[:, 1])
        st.eps = float(self.epsilon)
        self._fill_transforms(st)
        return st

    def forward_kinematics_batch(self, q, *, jacobian=False, layout="auto"):
        """q CUDA tensor [N, nq] (float64, or float32 for the FP32 kernels) -> T [N,4,4] (and
        J [N,6,nq]) of the same dtype."""
        _lib.require_cuda(q, "q")
        nq, dt = self._nq, _dtype_of(q)
        if layout == "auto":
            layout = _batch.detect_layout(q, nq)
        qs, ld = _batch.prep_input(q, nq, layout, dt)
        n, dev = qs.shape[0], qs.device
        T = _batch.alloc_view(n, (4, 4), layout, dt, dev)
        J = _batch.alloc_view(n, (6, nq), layout, dt, dev) if jacobian else None
        launch(getattr(_lib.load(), self._fk_fn + _SFX[dt]), dev, self._cached_struct(), _lib.ptr(qs), n, ld,
               _batch.flags_of(layout), _lib.ptr(T), _lib.ptr(J))
        return (T, J) if jacobian else T

    def inverse_kinematics_batch(self, T, branches=None, *, layout="auto"):
        """T CUDA tensor [N,4,4] -> (q [N,B,nq], valid [N,B]); branch order = `branches`
        (default: the canonical order, m outer, k inner)."""
        _lib.require_cuda(T, "T")
        branches = self._canonical if branches is None else branches
        n, dt = T.shape[0], _dtype_of(T)
        T2 = T.reshape(n, 16) if T.dim() == 3 else T
        if layout == "auto":
            layout = _batch.detect_layout(T2, 16)
        Ts, ld = _batch.prep_input(T2, 16, layout, dt)
        dev, nb, nq = Ts.device, len(branches), self._nq
        flat = []
        for b in branches:
            flat.extend([int(v) for v in (b if isinstance(b, (tuple, list)) else (b,))])
        br = (C.c_int8 * len(flat))(*flat)
        qv = _batch.alloc_view(n, (nb, nq), layout, dt, dev)
        vv = _batch.alloc_view(n, (nb,), layout, torch.uint8, dev)
        launch(getattr(_lib.load(), self._ik_fn + _SFX[dt]), dev, self._cached_struct(), _lib.ptr(Ts), n, ld,
               _batch.flags_of(layout), br, nb, _lib.ptr(qv), _lib.ptr(vv))
        return qv, vv.view(torch.bool)

    def forward_kinematics(self, q_values, plot=True):
        """Single pose -> 4x4 np.ndarray (1-element GPU batch; `plot` is ignored)."""
        q = torch.tensor(np.asarray(q_values, dtype=np.float64).reshape(1, self._nq), device="cuda")
        return self.forward_kinematics_batch(q, layout="aos")[0].cpu().numpy()

    def _ik_single(self, T, branch):
        Tt = torch.tensor(np.array(T, dtype=np.float64).reshape(1, 4, 4), device="cuda")
        q, valid = self.inverse_kinematics_batch(Tt, [branch], layout="aos")
        if not bool(valid[0, 0]):
            print("[INFO] The configuration is not reachable or a joint is out of limits")
        return q[0, 0].cpu().numpy()
