"""TrapezoidalGenerator: drop-in for utils/trapezoidal_generator.py of the reference.

Same stateful API (set_limits / set_frequency / generate_coefficients /
set_current_limits / set_positions / get_t / get_q / get_dq / get_ddq).  Sampling runs on
the GPU (rk_trap_sample_f64, one warp per profile) and is bit-identical to the
reference's NumPy output, including its shorter-than-n sample counts.  The batched entry
points are `generate_coefficients_batch` and `sample_batch`.
"""
import ctypes as C
from math import ceil

import numpy as np
import torch

from .. import _lib


class TrapezoidalGenerator():
    """Trapezoidal velocity profiles under velocity / acceleration limits."""

    def __init__(self, dq_max, ddq_max, control_freq=0):
        self._dq_max = dq_max
        self._ddq_max = ddq_max
        self._control_freq = control_freq

    def set_limits(self, dq_max, ddq_max):
        self._dq_max = dq_max
        self._ddq_max = ddq_max

    def set_frequency(self, control_freq):
        self._control_freq = control_freq

    def generate_coefficients(self, q_init, q_final):
        """t_1 (end of acceleration) and tau (start of deceleration) of one scalar move
        (trapezoidal_generator.py:36-79).  Host scalar arithmetic like the reference: this
        sets the generator's state; `generate_coefficients_batch` is the batched form."""
        self.set_positions(q_init, q_final)
        delta_q = abs(self._q_final - self._q_init)
        dq_max_prime = np.sqrt(delta_q * self._ddq_max)
        if dq_max_prime < self._dq_max:      # triangular profile
            t_1 = dq_max_prime / self._ddq_max
            tau = t_1
            self._dq_max_current = dq_max_prime
        else:                                # trapezoidal profile
            t_1 = self._dq_max / self._ddq_max
            tau = delta_q / self._dq_max
            self._dq_max_current = self._dq_max
        self._ddq_max_current = self._ddq_max
        if self._control_freq != 0:          # snap to the controller period
            t_1 = ceil(t_1 * self._control_freq) / self._control_freq
            tau = ceil(tau * self._control_freq) / self._control_freq
        return t_1, tau

    def set_current_limits(self, dq_max_current, ddq_max_current):
        self._dq_max_current = dq_max_current
        self._ddq_max_current = ddq_max_current

    def set_positions(self, q_init, q_final):
        self._q_init = q_init
        self._q_final = q_final

    @property
    def dq_max_current(self):
        return self._dq_max_current

    # ------------------------------------------------------------------ batched (GPU)
    def generate_coefficients_batch(self, q_init, q_final):
        """(t_1, tau) for N scalar moves; CUDA float64 tensors of any (equal) shape."""
        _lib.require_cuda(q_init, "q_init")
        a = q_init.to(torch.float64).contiguous()
        b = q_final.to(torch.float64).contiguous()
        t1, tau = torch.empty_like(a), torch.empty_like(a)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().rk_trap_coefficients_f64(
                _lib.ptr(a), _lib.ptr(b), a.numel(), float(self._dq_max), float(self._ddq_max),
                float(self._control_freq), _lib.ptr(t1), _lib.ptr(tau), _lib.stream_ptr(a.device)))
        return t1, tau

    @staticmethod
    def sample_batch(q_init, dq_cur, ddq_cur, t_1, tau, n):
        """get_q / get_dq / get_ddq / get_t for N profiles with explicit state (CUDA
        float64 tensors [N]).  Returns q, dq, ddq, t [N, n] (zero padded) and length [N]
        (int32; 0 where the reference raises for lack of samples)."""
        _lib.require_cuda(q_init, "q_init")
        args = [x.to(torch.float64).contiguous() for x in (q_init, dq_cur, ddq_cur, t_1, tau)]
        N, dev = args[0].numel(), args[0].device
        outs = [torch.empty((N, n), dtype=torch.float64, device=dev) for _ in range(4)]
        length = torch.empty(N, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().rk_trap_sample_f64(
                *[_lib.ptr(x) for x in args], N, int(n), *[_lib.ptr(o) for o in outs],
                _lib.ptr(length), _lib.stream_ptr(dev)))
        return (*outs, length)

    # ------------------------------------------------------------- reference methods
    def _sample(self, t_1, tau, n):
        f = lambda v: torch.tensor([float(v)], dtype=torch.float64, device="cuda")  # noqa: E731
        q, dq, ddq, t, length = self.sample_batch(
            f(self._q_init), f(self._dq_max_current), f(self._ddq_max_current), f(t_1), f(tau), n)
        L = int(length[0])
        if L == 0:
            raise ValueError("Not enough points. Please, increase n")
        return (x[0, :L].cpu().numpy() for x in (q, dq, ddq, t))

    @staticmethod
    def _flat_only(t_1, tau, n):
        """Number of constant-velocity samples when the acceleration segment is EMPTY
        (int(step * t_1) == 0, trapezoidal_generator.py:222-226), else None.  get_q / get_dq raise in
        that case (:142, :180) but get_t / get_ddq do not: they return the flat segment alone."""
        step = n / (tau + t_1)
        if int(step * t_1) != 0:
            return None
        return int(step * (tau - t_1))

    def get_t(self, t_1, tau, n):
        """Time samples (trapezoidal_generator.py:110-123)."""
        Lf = self._flat_only(t_1, tau, n)
        if Lf is not None:
            return np.linspace(t_1, tau, Lf)
        return list(self._sample(t_1, tau, n))[3]

    def get_q(self, t_1, tau, n):
        """Position samples (:125-159)."""
        return list(self._sample(t_1, tau, n))[0]

    def get_dq(self, t_1, tau, n):
        """Velocity samples (:161-186)."""
        return list(self._sample(t_1, tau, n))[1]

    def get_ddq(self, t_1, tau, n):
        """Acceleration samples (:188-206)."""
        Lf = self._flat_only(t_1, tau, n)
        if Lf is not None:
            return np.zeros(Lf)
        return list(self._sample(t_1, tau, n))[2]
