"""TrajectoryGenerator: drop-in for utils/trajectory_generator.py of the reference.

Single-trajectory methods keep the reference's signatures and return types (they run a
1-trajectory GPU batch); `generate_p2p_trajectory_batch`, `generate_lin_trajectory_batch`
and `get_dq_from_dx_batch` are the data-parallel entry points (one warp per trajectory).
Plotting arguments are accepted and ignored.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .polynomial_generator import PolynomialGenerator
from .trapezoidal_generator import TrapezoidalGenerator


class TrajectoryGenerator():
    """Polynomial, joint-space p2p and Cartesian linear trajectories."""

    def __init__(self, dq_max, ddq_max, dx_max, ddx_max, control_freq=0):
        self._pg = PolynomialGenerator()
        self._tg = TrapezoidalGenerator(dq_max=dq_max, ddq_max=ddq_max)
        self._dq_max = dq_max
        self._ddq_max = ddq_max
        self._dx_max = dx_max
        self._ddx_max = ddx_max
        self._control_freq = control_freq

    # ------------------------------------------------------------------ batched (GPU)
    def _p2p(self, qs_0, qs_f, n, v_max, a_max, robot=None, fk=False):
        _lib.require_cuda(qs_0, "qs_0")
        a = qs_0.to(torch.float64).contiguous()
        b = qs_f.to(torch.float64).contiguous()
        N, J = a.shape
        dev = a.device
        q, dq, ddq = (torch.empty((N, n, J), dtype=torch.float64, device=dev) for _ in range(3))
        t = torch.empty((N, n), dtype=torch.float64, device=dev)
        length = torch.empty(N, dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            if fk:
                if J != 3:
                    raise ValueError("fused FK needs the 3-joint RRR robot")
                T = torch.empty((N, n, 4, 4), dtype=torch.float64, device=dev)
                Jm = torch.empty((N, n, 6, 3), dtype=torch.float64, device=dev)
                _lib.check(lib.rk_rrr_p2p_fk_f64(
                    C.byref(robot._struct()), _lib.ptr(a), _lib.ptr(b), N, int(n), float(v_max),
                    float(a_max), float(self._control_freq), _lib.ptr(q), _lib.ptr(dq),
                    _lib.ptr(ddq), _lib.ptr(t), _lib.ptr(length), _lib.ptr(T), _lib.ptr(Jm),
                    _lib.stream_ptr(dev)))
                return q, dq, ddq, t, length, T, Jm
            _lib.check(lib.rk_trap_p2p_f64(
                _lib.ptr(a), _lib.ptr(b), N, J, int(n), float(v_max), float(a_max),
                float(self._control_freq), _lib.ptr(q), _lib.ptr(dq), _lib.ptr(ddq), _lib.ptr(t),
                _lib.ptr(length), _lib.stream_ptr(dev)))
        return q, dq, ddq, t, length

    def generate_p2p_trajectory_batch(self, qs_0, qs_f, n=100, robot=None):
        """Equalised trapezoidal joint trajectories for N (start, goal) pairs.

        qs_0, qs_f: CUDA tensors [N, J], J <= 8.  Returns q, dq, ddq [N, n, J], t [N, n]
        and length [N]: the reference produces FEWER than n samples; rows >= length[i] are
        zero padding and length[i] == 0 marks a trajectory for which the reference raises.
        With robot=RRRRobot(...) the per-sample forward kinematics and 6x3 Jacobian are
        fused into the same launch and T [N,n,4,4], J [N,n,6,3] are returned as well."""
        return self._p2p(qs_0, qs_f, n, self._dq_max, self._ddq_max, robot, fk=robot is not None)

    def generate_lin_trajectory_batch(self, p_0, p_f, n=100, robot=None):
        """Cartesian straight-line trajectories for N (p_0, p_f) pairs ([N,3] CUDA tensors).
        Without a robot: x, dx, ddx [N,n,3], t, length.  With robot=RRRRobot(...): per
        sample inverse kinematics (m = k = 1, as RRRRobot.move_via_points) and joint
        velocities dq = J_v(q)^-1 dx (get_dq_from_dx): x, dx, t, q, dq, valid, length."""
        if robot is None:
            return self._p2p(p_0, p_f, n, self._dx_max, self._ddx_max)
        _lib.require_cuda(p_0, "p_0")
        a = p_0.to(torch.float64).contiguous()
        b = p_f.to(torch.float64).contiguous()
        N, dev = a.shape[0], a.device
        x, dx, qj, dqj = (torch.empty((N, n, 3), dtype=torch.float64, device=dev) for _ in range(4))
        t = torch.empty((N, n), dtype=torch.float64, device=dev)
        valid = torch.empty((N, n), dtype=torch.uint8, device=dev)
        length = torch.empty(N, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().rk_rrr_lin_ik_f64(
                C.byref(robot._struct()), _lib.ptr(a), _lib.ptr(b), N, int(n), float(self._dx_max),
                float(self._ddx_max), float(self._control_freq), _lib.ptr(x), _lib.ptr(dx),
                _lib.ptr(t), _lib.ptr(qj), _lib.ptr(dqj), _lib.ptr(valid), _lib.ptr(length),
                _lib.stream_ptr(dev)))
        return x, dx, t, qj, dqj, valid.bool(), length

    def generate_joint_poly_trajectory_batch(self, qs_0, qs_f, t_f, t_0=0.0, n=100):
        """Quintic joint trajectories for N x J boundary conditions: qs_0, qs_f CUDA tensors
        [N, J, 3] -> q, dq, ddq [N, n, J] (generate_joint_poly_trajectory, :27-69, batched)."""
        N, J = qs_0.shape[0], qs_0.shape[1]
        _, q, dq, ddq = self._pg.profiles_batch(qs_0.reshape(N * J, 3), qs_f.reshape(N * J, 3),
                                                t_f, t_0, n)
        shape = lambda x: x.reshape(N, J, n).permute(0, 2, 1)  # noqa: E731
        return shape(q), shape(dq), shape(ddq)

    @staticmethod
    def get_dq_from_dx_batch(qs, dxs, robot):
        """dq = J_v(q)^-1 dx for [N,3] CUDA tensors through the RRR Jacobian kernel."""
        bare = type(robot)()
        _, J = bare.forward_kinematics_batch(qs, jacobian=True, layout="aos")
        return torch.linalg.solve(J[:, :3, :3], dxs.to(torch.float64).unsqueeze(-1)).squeeze(-1)

    # ------------------------------------------------------------- reference methods
    def generate_joint_poly_trajectory(self, qs_0, qs_f, t_f, t_0=0.0, n=100, plot=True):
        """Quintic joint trajectory (trajectory_generator.py:27-69), host NumPy."""
        coefs = [self._pg.generate_coefficients(q_0, q_f, t_f, t_0) for q_0, q_f in zip(qs_0, qs_f)]
        return np.array([self._pg.polynomial_from_coefs(c, t_0, t_f, n) for c in coefs]).T

    def _single(self, starts, goals, n, v_max, a_max):
        a = torch.tensor(np.asarray(starts, dtype=np.float64).reshape(1, -1), device="cuda")
        b = torch.tensor(np.asarray(goals, dtype=np.float64).reshape(1, -1), device="cuda")
        q, dq, ddq, t, length = self._p2p(a, b, n, v_max, a_max)
        L = int(length[0])
        if L == 0:
            raise ValueError("Not enough points. Please, increase n")
        return (q[0, :L].cpu().numpy(), dq[0, :L].cpu().numpy(), ddq[0, :L].cpu().numpy(),
                t[0, :L].cpu().numpy())

    def generate_p2p_trajectory(self, qs_0, qs_f, n=100, plot=False):
        """trajectory_generator.py:71-92: array [L, J] of joint positions, L <= n."""
        self._tg.set_frequency(self._control_freq)
        self._tg.set_limits(self._dq_max, self._ddq_max)
        return self._single(qs_0, qs_f, n, self._dq_max, self._ddq_max)[0]

    def generate_lin_trajectory(self, p_0, p_f, n=100, plot=False):
        """trajectory_generator.py:94-116: (xs [L,3], dxs [L,3], ts [L])."""
        self._tg.set_frequency(self._control_freq)
        self._tg.set_limits(self._dx_max, self._ddx_max)
        xs, dxs, _, ts = self._single(p_0, p_f, n, self._dx_max, self._ddx_max)
        return xs, dxs, ts

    def _generate_equalized_profiles(self, qs_0, qs_f, n):
        """trajectory_generator.py:118-159 under the generator's current limits: lists of
        per-joint arrays qs, dqs, ddqs and the time vector."""
        q, dq, ddq, t = self._single(qs_0, qs_f, n, self._tg._dq_max, self._tg._ddq_max)
        return list(q.T), list(dq.T), list(ddq.T), t

    def get_dq_from_dx(self, qs, dxs, J_inv=None, ts=None):
        """trajectory_generator.py:161-188.  qs is [J, L] (as returned by
        RRRRobot.move_via_points), dxs is [L, 3]; returns dqs [J, L].  J_inv (the
        reference's symbolic inverse) is accepted and ignored: the numeric RRR Jacobian
        kernel is used instead."""
        from ..robots.rrr_robot import RRRRobot
        q = torch.tensor(np.asarray(qs, dtype=np.float64).T.copy(), device="cuda")
        dx = torch.tensor(np.asarray(dxs, dtype=np.float64), device="cuda")
        return self.get_dq_from_dx_batch(q, dx, RRRRobot()).cpu().numpy().T

    def interpolate(self, p_start, p_finish, n):
        """Linear interpolation between two points (trajectory_generator.py:190-204)."""
        v = np.array([p_finish]) - np.array([p_start])
        t = np.array([np.linspace(0, 1, n)]).T
        return p_start + t.dot(v)
