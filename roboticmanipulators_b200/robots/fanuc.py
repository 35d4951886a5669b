"""Fanuc165F: drop-in for robots/fanuc.py of the reference, with batched B200 kernels.

Same constructor, attributes and single-pose methods as the reference
(robots/fanuc.py:16-302); the single-pose methods run a 1-element batch through the
same CUDA kernels (there is no CPU path).  New: `forward_kinematics_batch`,
`inverse_kinematics_batch`, `jacobian_batch`.
"""
import ctypes as C

import numpy as np
import torch

from .. import _batch, _lib
from .robot import Robot, launch

#: the canonical 8 branches, in the order every batched IK call returns them
CANONICAL_8 = tuple((m, k, w) for m in (1, -1) for k in (1, -1) for w in (0, 1))
#: all 12 combinations the reference's signature admits (w in {0, 1, -1})
ALL_12 = tuple((m, k, w) for m in (1, -1) for k in (1, -1) for w in (0, 1, -1))

_DTYPES = {torch.float64: "f64", torch.float32: "f32"}
_HOST_PIPES = {}      # (device index, chunk, slots) -> RkHostPipe* (device staging ring + 3 streams)


class Fanuc165F(Robot):
    """Fanuc165F forward / inverse kinematics.

    Attributes mirror the reference: qs_lim_deg, qs_lim_rad, qs_offsets, qs_directions,
    d, dq, T_base, T_tool, epsilon (robots/fanuc.py:21-41)."""

    qs_lim_deg = ((-185.0, 185.0),
                  (-60.0, 76.0),
                  (-156.0, 156.0),
                  (-360.0, 360.0),
                  (-125.0, 125.0),
                  (-360.0, 360.0))

    def __init__(self, T_base=None, T_tool=None,
                 lengths=None, save=True, offsets=None, directions=None):
        """Same arguments as robots/fanuc.py:43-44.  `save` is accepted for
        compatibility: there is nothing to cache (the robot is a ~700-byte constant
        block passed to each launch, not pickled SymPy matrices)."""
        self.set_transforms(T_base, T_tool)
        self.set_lengths(lengths)
        self.set_joint_offsets(offsets)
        self.set_joint_directions(directions)
        self._generate_value_pairs()
        self._calculate_limits_radians()
        self._save = save

    def set_lengths(self, lengths):
        if lengths is None:
            self._ls = (346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0)  # fanuc.py:73
        else:
            self._ls = lengths

    def set_joint_offsets(self, offsets):
        if offsets is None:
            angle = np.pi / 2
            self.qs_offsets = np.array([0.0, -angle, angle, 0.0, 0.0, 0.0])  # fanuc.py:79-80
        else:
            self.qs_offsets = offsets

    def set_joint_directions(self, directions):
        if directions is None:
            self.qs_directions = np.array([1, 1, -1, -1, -1, -1])           # fanuc.py:86
        else:
            self.qs_directions = directions

    def _generate_value_pairs(self):
        """d, dq of the TzTx -> Ry Tx Ry^-1 substitution (fanuc.py:90-104)."""
        self._value_pairs = [(f"l_{i}", self._ls[i]) for i in range(len(self._ls))]
        self.d = np.sqrt(self._ls[4] * self._ls[4] + self._ls[5] * self._ls[5])
        self.dq = np.arctan2(self._ls[4], self._ls[5])
        self._value_pairs.append(("d", self.d))
        self._value_pairs.append(("dq", self.dq))

    def _calculate_limits_radians(self):
        self.qs_lim_rad = tuple(
            (np.deg2rad(x[0]), np.deg2rad(x[1])) for x in self.qs_lim_deg)   # fanuc.py:110-111

    # ------------------------------------------------------------------ C-ABI struct
    def _fingerprint(self):
        return (np.asarray(self.qs_offsets, dtype=np.float64).tobytes()
                + np.asarray(self.qs_directions, dtype=np.float64).tobytes()
                + np.asarray(self._ls, dtype=np.float64).tobytes())

    def _struct(self):
        st = _lib.RkFanuc()
        _lib.fill(st.l, self._ls)
        _lib.fill(st.offsets, self.qs_offsets)
        _lib.fill(st.directions, self.qs_directions)
        lim = np.asarray(self.qs_lim_rad, dtype=np.float64)
        _lib.fill(st.lim_lo, lim[:, 0])
        _lib.fill(st.lim_hi, lim[:, 1])
        st.d, st.dq, st.eps = float(self.d), float(self.dq), float(self.epsilon)
        self._fill_transforms(st)
        return st

    # ------------------------------------------------------------------ batched API
    def forward_kinematics_batch(self, q, *, jacobian=False, measure=False,
                                 layout="auto", user_angles=False):
        """FK (+ 6x6 Jacobian, + det J) for a batch of joint vectors.

        q: CUDA tensor, logical shape [N, 6], float64 or float32.  A transposed view
        of a contiguous [6, N] tensor is consumed as SOA without a copy.
        Returns T [N,4,4]; with jacobian=True also J [N,6,6] (columns w.r.t. the
        model angles q' = q*dir + off, the convention of fanuc_jacobians.py, unless
        user_angles=True); with measure=True also det J [N] (|det| = manipulability).
        """
        _lib.require_cuda(q, "q")
        if q.dtype not in _DTYPES:
            raise TypeError("q must be float64 or float32")
        if layout == "auto":
            layout = _batch.detect_layout(q, 6)
        qs, ld = _batch.prep_input(q, 6, layout, q.dtype)
        n = qs.shape[0]
        dev = qs.device
        T = _batch.alloc_view(n, (4, 4), layout, q.dtype, dev)
        J = det = None
        if jacobian:
            J = _batch.alloc_view(n, (6, 6), layout, q.dtype, dev)
        if measure:
            det = torch.empty(n, dtype=q.dtype, device=dev)
        flags = _batch.flags_of(layout) | (_lib.RK_JAC_USER_ANGLES if user_angles else 0)
        fn = getattr(_lib.load(), "rk_fanuc_fk_jac_" + _DTYPES[q.dtype])
        launch(fn, dev, self._cached_struct(), qs.data_ptr(), n, ld, flags, T.data_ptr(), _lib.ptr(J),
               _lib.ptr(det))
        out = (T,)
        if jacobian:
            out += (J,)
        if measure:
            out += (det,)
        return out if len(out) > 1 else T

    def jacobian_batch(self, q, *, layout="auto", user_angles=False):
        """6x6 Jacobian and det J only (no pose traffic)."""
        _lib.require_cuda(q, "q")
        if q.dtype not in _DTYPES:
            raise TypeError("q must be float64 or float32")
        if layout == "auto":
            layout = _batch.detect_layout(q, 6)
        qs, ld = _batch.prep_input(q, 6, layout, q.dtype)
        n, dev = qs.shape[0], qs.device
        J = _batch.alloc_view(n, (6, 6), layout, q.dtype, dev)
        det = torch.empty(n, dtype=q.dtype, device=dev)
        flags = _batch.flags_of(layout) | (_lib.RK_JAC_USER_ANGLES if user_angles else 0)
        fn = getattr(_lib.load(), "rk_fanuc_fk_jac_" + _DTYPES[q.dtype])
        launch(fn, dev, self._cached_struct(), qs.data_ptr(), n, ld, flags, None, J.data_ptr(),
               det.data_ptr())
        return J, det

    def inverse_kinematics_batch(self, T, branches=CANONICAL_8, *, layout="auto"):
        """Pieper IK over `branches` (list of (m, k, w)) for a batch of poses.

        T: CUDA tensor [N,4,4] (a permuted view of a contiguous [16, N] is SOA).
        Returns (q [N, B, 6], valid [N, B] bool); branch order = `branches`; failed
        branches hold the reference's all-zero sentinel."""
        _lib.require_cuda(T, "T")
        if T.dtype not in _DTYPES:
            raise TypeError("T must be float64 or float32")
        n = T.shape[0]
        T2 = T.reshape(n, 16) if T.dim() == 3 else T
        if layout == "auto":
            layout = _batch.detect_layout(T2, 16)
        Ts, ld = _batch.prep_input(T2, 16, layout, T.dtype)
        dev = Ts.device
        nb = len(branches)
        # NULL selects the canonical 8 in the library (and skips building the ctypes array)
        br = None if branches is CANONICAL_8 else (C.c_int8 * (3 * nb))(*[int(v) for b in branches for v in b])
        qv = _batch.alloc_view(n, (nb, 6), layout, T.dtype, dev)
        vv = _batch.alloc_view(n, (nb,), layout, torch.uint8, dev)
        fn = getattr(_lib.load(), "rk_fanuc_ik_" + _DTYPES[T.dtype])
        launch(fn, dev, self._cached_struct(), Ts.data_ptr(), n, ld, _batch.flags_of(layout), br, nb,
               qv.data_ptr(), vv.data_ptr())
        return qv, vv.view(torch.bool)

    def fk_jac_ik_batch(self, q, *, layout="auto", user_angles=False):
        """forward_kinematics_batch(q, jacobian=True, measure=True) followed by
        inverse_kinematics_batch(T) over the canonical 8 branches, as one call (one kernel for
        SOA inputs: the pose does not travel through memory between the two halves; the FK -> IK
        round trip of fanuc_kinematics.py / tests/test_robots.py:33-52).
        Returns (T [N,4,4], J [N,6,6], det [N], q_ik [N,8,6], valid [N,8] bool)."""
        _lib.require_cuda(q, "q")
        if q.dtype not in _DTYPES:
            raise TypeError("q must be float64 or float32")
        if layout == "auto":
            layout = _batch.detect_layout(q, 6)
        qs, ld = _batch.prep_input(q, 6, layout, q.dtype)
        n, dev = qs.shape[0], qs.device
        dt = q.dtype
        T = _batch.alloc_view(n, (4, 4), layout, dt, dev)
        J = _batch.alloc_view(n, (6, 6), layout, dt, dev)
        det = torch.empty(n, dtype=dt, device=dev)
        qv = _batch.alloc_view(n, (8, 6), layout, dt, dev)
        vv = _batch.alloc_view(n, (8,), layout, torch.uint8, dev)
        flags = _batch.flags_of(layout) | (_lib.RK_JAC_USER_ANGLES if user_angles else 0)
        fn = getattr(_lib.load(), "rk_fanuc_fk_jac_ik8_" + _DTYPES[q.dtype])
        launch(fn, dev, self._cached_struct(), qs.data_ptr(), n, ld, flags, T.data_ptr(), J.data_ptr(),
               det.data_ptr(), qv.data_ptr(), vv.data_ptr())
        return T, J, det, qv, vv.view(torch.bool)

    def dq_from_dx_batch(self, q, dx, *, layout="aos", user_angles=True):
        """Joint rates from Cartesian twists: dq = J(q)^-1 dx for N poses, solved in the kernel that
        computes the 6x6 Jacobian (TrajectoryGenerator.get_dq_from_dx, utils/trajectory_generator.py:
        161-188, for this robot).  q, dx: CUDA [N, 6] (dx = [v; omega]); user_angles=True returns
        rates of the USER joint angles (columns scaled by qs_directions).  Returns (dq [N, 6],
        ok [N] bool -- False where J is exactly singular)."""
        _lib.require_cuda(q, "q")
        _lib.require_cuda(dx, "dx")
        if q.dtype not in _DTYPES:
            raise TypeError("q must be float64 or float32")
        if layout not in ("aos", "soa"):
            raise ValueError("layout must be 'aos' or 'soa'")
        qs, ld = _batch.prep_input(q, 6, layout, q.dtype)
        ds, _ = _batch.prep_input(dx.reshape(dx.shape[0], 6), 6, layout, q.dtype)
        n, dev = qs.shape[0], qs.device
        dq = _batch.alloc_view(n, (6,), layout, q.dtype, dev)
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        flags = _batch.flags_of(layout) | (_lib.RK_JAC_USER_ANGLES if user_angles else 0)
        fn = getattr(_lib.load(), "rk_fanuc_dq_from_dx_" + _DTYPES[q.dtype])
        launch(fn, dev, self._cached_struct(), _lib.ptr(qs), _lib.ptr(ds), n, ld, flags, dq.data_ptr(),
               _lib.ptr(ok))
        return dq, ok.view(torch.bool)

    _STREAM_OUT = {"T": (4, 4), "J": (6, 6), "det": (), "q_ik": (8, 6), "valid": (8,)}

    def fk_jac_ik_stream(self, q, *, want=("T", "J", "det", "q_ik", "valid"), out=None,
                         chunk=1 << 16, slots=3, layout="aos", user_angles=False, device=None):
        """The round trip of fk_jac_ik_batch on HOST arrays (NumPy arrays or CPU torch tensors, as
        the reference keeps them), streamed through the GPU: the batch is cut in chunks that are
        uploaded, computed and downloaded on three streams, so both directions of the host link
        are busy all the time -- the link, not the GPU, is the roofline of this call (48 B up and
        up to 816 B down per pose in float64).

        q: [N, 6] float64 / float32 on the host (page-locked memory -- `torch.empty(...,
        pin_memory=True)` -- gives full link speed).  `want` selects what is computed AND shipped
        (e.g. ("q_ik", "valid") moves 392 B per pose instead of 816).  `out` may hold preallocated
        (pinned) result tensors by name; missing ones are allocated page-locked.  layout="soa":
        q is a [N, 6] view of [6, N] planes and results are such views too.
        Returns a dict name -> host tensor with the reference's shapes (T [N,4,4], J [N,6,6],
        det [N], q_ik [N,8,6], valid [N,8] bool); complete when the call returns."""
        if isinstance(q, np.ndarray):
            q = torch.from_numpy(q)
        if not isinstance(q, torch.Tensor) or q.is_cuda:
            raise _lib.RkError("fk_jac_ik_stream takes HOST arrays; use fk_jac_ik_batch for CUDA tensors")
        if q.dtype not in _DTYPES:
            raise TypeError("q must be float64 or float32")
        unknown = set(want) - set(self._STREAM_OUT)
        if unknown or not want:
            raise ValueError(f"want must name some of {sorted(self._STREAM_OUT)}")
        n = q.shape[0]
        soa = layout == "soa"
        if layout not in ("aos", "soa"):
            raise ValueError("layout must be 'aos' or 'soa'")
        q = q.reshape(n, 6)
        if soa:
            if not (q.stride(0) == 1 and (n == 1 or q.stride(1) >= n)):
                q = q.t().contiguous().t()
            ld = q.stride(1) if n > 1 else 1
        else:
            q, ld = q.contiguous(), n
        names = set(want)
        if names & {"q_ik", "valid"}:
            names |= {"q_ik", "valid"}
        res, ptrs = {}, {}
        for name, shape in self._STREAM_OUT.items():
            if name not in names:
                ptrs[name] = None
                continue
            dt = torch.uint8 if name == "valid" else q.dtype
            t = None if out is None else out.get(name)
            ncomp = int(np.prod(shape)) if shape else 1
            if t is not None:
                if t.dtype == torch.bool:
                    t = t.view(torch.uint8)
                ok = (t.dtype == dt and not t.is_cuda and t.numel() == n * ncomp and
                      (t.is_contiguous() if not soa else True))
                if not ok or soa:
                    raise ValueError(f"out[{name!r}] must be a contiguous host tensor of {n}x{shape} {dt} (AOS)")
                store = t
            elif soa:
                store = torch.empty((ncomp, ld), dtype=dt, pin_memory=True)
            else:
                store = torch.empty((n, *shape), dtype=dt, pin_memory=True)
            ptrs[name] = C.c_void_p(store.data_ptr())
            if soa:
                view = store[:, :n].view(*shape, n) if shape else store[0, :n]
                view = view.permute(len(shape), *range(len(shape))) if shape else view
            else:
                view = store.view(n, *shape) if shape else store.view(n)
            res[name] = view.view(torch.bool) if name == "valid" else view
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        pipe = self._host_pipe(dev, int(chunk), int(slots))
        flags = (_lib.RK_SOA if soa else _lib.RK_AOS) | (_lib.RK_JAC_USER_ANGLES if user_angles else 0)
        fn = getattr(_lib.load(), "rk_fanuc_fk_jac_ik8_host_" + _DTYPES[q.dtype])
        with torch.cuda.device(dev):
            _lib.check(fn(pipe, self._cached_struct(), C.c_void_p(q.data_ptr()), n, ld, flags,
                          ptrs["T"], ptrs["J"], ptrs["det"], ptrs["q_ik"], ptrs["valid"]))
        return {k: res[k] for k in want}

    def inverse_kinematics_stream(self, T, *, out=None, chunk=1 << 16, slots=3, device=None):
        """inverse_kinematics_batch (the 8 canonical branches) on HOST arrays: targets T [N,4,4]
        (NumPy or CPU torch, float64 / float32; page-locked for full link speed) are streamed
        through the GPU in chunks on three streams -- 128 B up and 392 B down per target in
        float64.  `out` may hold preallocated (pinned) "q_ik" [N,8,6] / "valid" [N,8] tensors.
        Returns (q_ik, valid) host tensors, complete when the call returns."""
        if isinstance(T, np.ndarray):
            T = torch.from_numpy(T)
        if not isinstance(T, torch.Tensor) or T.is_cuda:
            raise _lib.RkError("inverse_kinematics_stream takes HOST arrays; use inverse_kinematics_batch "
                               "for CUDA tensors")
        if T.dtype not in _DTYPES:
            raise TypeError("T must be float64 or float32")
        n = T.shape[0]
        T = T.reshape(n, 16).contiguous()
        res = {}
        for name, shape, dt in (("q_ik", (n, 8, 6), T.dtype), ("valid", (n, 8), torch.uint8)):
            t = None if out is None else out.get(name)
            if t is not None:
                if t.dtype == torch.bool:
                    t = t.view(torch.uint8)
                if t.dtype != dt or t.is_cuda or tuple(t.shape) != shape or not t.is_contiguous():
                    raise ValueError(f"out[{name!r}] must be a contiguous host tensor of shape {shape}, {dt}")
            else:
                t = torch.empty(shape, dtype=dt, pin_memory=True)
            res[name] = t
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        pipe = self._host_pipe(dev, int(chunk), int(slots))
        fn = getattr(_lib.load(), "rk_fanuc_ik8_host_" + _DTYPES[T.dtype])
        with torch.cuda.device(dev):
            _lib.check(fn(pipe, self._cached_struct(), C.c_void_p(T.data_ptr()), n, n, _lib.RK_AOS,
                          C.c_void_p(res["q_ik"].data_ptr()), C.c_void_p(res["valid"].data_ptr())))
        return res["q_ik"], res["valid"].view(torch.bool)

    def _host_pipe(self, dev, chunk, slots):
        """The (device, chunk, slots) staging pipe of this process, created on first use."""
        key = (dev.index, chunk, slots)
        pipe = _HOST_PIPES.get(key)
        if pipe is None:
            pipe = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(_lib.load().rk_host_pipe_create(C.byref(pipe), chunk, slots))
            _HOST_PIPES[key] = pipe
        return pipe

    # ------------------------------------------------------------- reference methods
    def forward_kinematics(self, q_values, plot=True):
        """robots/fanuc.py:151-179.  Returns a 4x4 np.ndarray.  `plot` is accepted and
        ignored (plotting is outside the accelerated path)."""
        q = torch.tensor(np.asarray(q_values, dtype=np.float64).reshape(1, 6), device="cuda")
        T = self.forward_kinematics_batch(q, layout="aos")
        return T[0].cpu().numpy()

    def inverse_kinematics(self, T, m=1, k=1, w=0):
        """robots/fanuc.py:196-302.  Returns np.ndarray[6], zeros on failure."""
        if abs(m) != 1:
            print("[WARNING] m can only be -1 or 1. Defaulting to 1")
            m = 1
        if abs(k) != 1:
            print("[WARNING] k can only be -1 or 1. Defaulting to 1")
            k = 1
        if abs(w) != 1 and w != 0:
            print("[WARNING] w can only be -1, 0 or 1. Defaulting to 1")
            w = 1
        Tt = torch.tensor(np.array(T, dtype=np.float64).reshape(1, 4, 4), device="cuda")
        q, valid = self.inverse_kinematics_batch(Tt, [(m, k, w)], layout="aos")
        if not bool(valid[0, 0]):
            print("[INFO] The configuration is not reachable or a joint is out of limits")
        return q[0, 0].cpu().numpy()
