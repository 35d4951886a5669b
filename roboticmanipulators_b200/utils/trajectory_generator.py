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
    def _p2p(self, qs_0, qs_f, n, v_max, a_max, robot=None, fk=False, dtype=torch.float64):
        _lib.require_cuda(qs_0, "qs_0")
        a = qs_0.to(torch.float64).contiguous()
        b = qs_f.to(torch.float64).contiguous()
        N, J = a.shape
        dev = a.device
        if dtype not in (torch.float64, torch.float32) or (dtype == torch.float32 and not fk):
            raise ValueError("dtype=torch.float32 is available for the fused FK / Jacobian sampler only")
        q, dq, ddq = (torch.empty((N, n, J), dtype=dtype, device=dev) for _ in range(3))
        t = torch.empty((N, n), dtype=dtype, device=dev)
        length = torch.empty(N, dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            if fk:
                if J != 3:
                    raise ValueError("fused FK needs the 3-joint RRR robot")
                T = torch.empty((N, n, 4, 4), dtype=dtype, device=dev)
                Jm = torch.empty((N, n, 6, 3), dtype=dtype, device=dev)
                fn = lib.rk_rrr_p2p_fk_f64 if dtype == torch.float64 else lib.rk_rrr_p2p_fk_f32
                _lib.check(fn(
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

    def generate_p2p_trajectory_batch(self, qs_0, qs_f, n=100, robot=None, dtype=torch.float64):
        """Equalised trapezoidal joint trajectories for N (start, goal) pairs.

        qs_0, qs_f: CUDA tensors [N, J], J <= 8.  Returns q, dq, ddq [N, n, J], t [N, n]
        and length [N]: the reference produces FEWER than n samples; rows >= length[i] are
        zero padding and length[i] == 0 marks a trajectory for which the reference raises.
        With robot=RRRRobot(...) the per-sample forward kinematics and 6x3 Jacobian are
        fused into the same launch and T [N,n,4,4], J [N,n,6,3] are returned as well; with
        dtype=torch.float32 the outputs and that per-sample FK are FP32 while the profile (hence
        `length`) is still evaluated in FP64, bit-identical to the FP64 call."""
        return self._p2p(qs_0, qs_f, n, self._dq_max, self._ddq_max, robot, fk=robot is not None,
                         dtype=dtype)

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
        """dq = J(q)^-1 dx for N samples, the Jacobian formed and the system solved inside one kernel
        (no library solve).  `robot` selects the Jacobian:
          * RRRRobot: the 3x3 position block of RzTzRyTxRyTx with the robot's link lengths and identity
            base / tool, as rrr_robot_jacobians.py:14-46 builds the reference's J_inv; qs, dxs [N,3];
          * Fanuc165F: the 6x6 Jacobian w.r.t. the user joint angles; qs [N,6], dxs [N,6] twists;
          * JacobianCalculator: any chain -- qs are the token values [N, ntok], dxs [N, n_active], the
            first n_active Jacobian rows form the system (see JacobianCalculator.dq_from_dx_batch).
        Returns dq (same shape as dxs)."""
        from ..robots.fanuc import Fanuc165F
        from ..robots.rrr_robot import RRRRobot
        from .jacobians import JacobianCalculator
        from .robo_math import chain_dq_from_dx
        if isinstance(robot, Fanuc165F):
            return robot.dq_from_dx_batch(qs, dxs)[0]
        if isinstance(robot, JacobianCalculator):
            return robot.dq_from_dx_batch(qs, dxs)[0]
        if isinstance(robot, RRRRobot):
            ls = [float(v) for v in robot.ls]
            return chain_dq_from_dx(["Rz", "Tz", "Ry", "Tx", "Ry", "Tx"], qs.to(torch.float64), dxs,
                                    [1, 0, 1, 0, 1, 0], rows=(0, 1, 2), src=[0, -1, 1, -1, 2, -1],
                                    cst=[0.0, ls[0], 0.0, ls[1], 0.0, ls[2]])[0]
        raise TypeError("robot must be an RRRRobot, a Fanuc165F or a JacobianCalculator")

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

    def get_dq_from_dx(self, qs, dxs, J_inv=None, ts=None, robot=None):
        """trajectory_generator.py:161-188.  qs is [J, L] (as returned by RRRRobot.move_via_points),
        dxs is [L, J]; returns dqs [J, L].

        The reference takes a SYMBOLIC inverse Jacobian J_inv in the symbols q_0 .. q_{J-1} and
        evaluates it per sample.  Here:
          * robot=<RRRRobot | Fanuc165F | JacobianCalculator>: the Jacobian is formed and the system
            solved on the GPU (get_dq_from_dx_batch) -- the accelerated path;
          * J_inv given (a sympy Matrix) and no robot: the caller's own expression is evaluated for
            all samples at once on the GPU (sympy.lambdify with torch as the numeric module over CUDA
            tensors: the formula is the caller's, there is no hand-written kernel for an arbitrary
            expression);
          * neither: the default RRRRobot() of rrr_robot_jacobians.py, on the GPU."""
        qs = np.asarray(qs, dtype=np.float64)
        dxs = np.asarray(dxs, dtype=np.float64)
        if robot is None and J_inv is not None:
            import sympy as sp
            Jm = sp.Matrix(J_inv)
            nq, L = qs.shape
            syms = [sp.symbols(f"q_{i}") for i in range(nq)]
            t = sp.symbols("t")        # a J_inv built with f_of_t=True carries q_i(t): same variables
            Jm = Jm.subs({sp.Function(f"q_{i}")(t): syms[i] for i in range(nq)})
            qd = torch.tensor(qs, device="cuda")                       # [J, L] on the GPU
            dxd = torch.tensor(dxs, device="cuda")
            rows, cols = Jm.shape
            Ji = torch.empty((L, rows, cols), dtype=torch.float64, device="cuda")
            for r in range(rows):
                for c in range(cols):
                    f = sp.lambdify(syms, Jm[r, c], modules="torch")   # the caller's own formula, on the GPU
                    Ji[:, r, c] = torch.as_tensor(f(*[qd[i] for i in range(nq)]), dtype=torch.float64,
                                                  device="cuda")
            return torch.einsum("lrc,lc->rl", Ji, dxd).cpu().numpy()
        if robot is None:
            from ..robots.rrr_robot import RRRRobot
            robot = RRRRobot()
        q = torch.tensor(qs.T.copy(), device="cuda")
        dx = torch.tensor(dxs, device="cuda")
        return self.get_dq_from_dx_batch(q, dx, robot).cpu().numpy().T

    def interpolate(self, p_start, p_finish, n):
        """Linear interpolation between two points (trajectory_generator.py:190-204)."""
        v = np.array([p_finish]) - np.array([p_start])
        t = np.array([np.linspace(0, 1, n)]).T
        return p_start + t.dot(v)
