"""Quintic polynomial profiles: drop-in for utils/polynomial_generator.py.  The scalar methods
are host NumPy exactly like the reference; `profiles_batch` evaluates N profiles (coefficients +
position / velocity / acceleration samples) in one CUDA launch (rk_poly5_f64)."""
from math import comb, factorial

import numpy as np


class PolynomialGenerator():
    def generate_coefficients(self, q_init, q_final, t_f, t_0=0.0):
        """Solve the 6x6 boundary-condition system for [q, dq, ddq] at t_0 and t_f
        (polynomial_generator.py:16-36)."""
        A = np.vstack((self._get_constraint_submatrix(t_0), self._get_constraint_submatrix(t_f)))
        rhs = np.hstack((np.array(q_init), np.array(q_final)))
        return np.linalg.inv(A).dot(rhs)

    def _get_constraint_submatrix(self, t):
        return np.array([[1, t, t**2, t**3, t**4, t**5],
                         [0, 1, 2 * t, 3 * t**2, 4 * t**3, 5 * t**4],
                         [0, 0, 2, 6 * t, 12 * t**2, 20 * t**3]])

    def polynomial_from_coefs(self, coefs, from_t, to_t, n):
        ts = np.linspace(from_t, to_t, n)
        return sum(c * ts**i for i, c in enumerate(coefs))

    def dpolynomial_from_coefs(self, d, coefs, from_t, to_t, n):
        ts = np.linspace(from_t, to_t, n)
        if d >= len(coefs):
            return np.zeros(n)
        return sum(c * factorial(d) * comb(i, d) * ts**(i - d)
                   for i, c in enumerate(coefs[d:], d))

    @staticmethod
    def profiles_batch(b_init, b_final, t_f, t_0=0.0, n=100):
        """N quintic profiles at once.  b_init, b_final: CUDA tensors [N, 3] of (q, dq, ddq) at
        t_0 and t_f (scalars or [N] tensors).  Returns coefs [N,6], q, dq, ddq [N,n]."""
        import torch
        from .. import _lib
        _lib.require_cuda(b_init, "b_init")
        a = b_init.to(torch.float64).contiguous()
        b = b_final.to(torch.float64).contiguous()
        N, dev = a.shape[0], a.device

        def vec(x):
            if isinstance(x, torch.Tensor):
                return x.to(device=dev, dtype=torch.float64).contiguous()
            return torch.full((N,), float(x), dtype=torch.float64, device=dev)
        t0, tf = vec(t_0), vec(t_f)
        coef = torch.empty((N, 6), dtype=torch.float64, device=dev)
        q, dq, ddq = (torch.empty((N, n), dtype=torch.float64, device=dev) for _ in range(3))
        with torch.cuda.device(dev):
            _lib.check(_lib.load().rk_poly5_f64(
                _lib.ptr(a), _lib.ptr(b), _lib.ptr(t0), _lib.ptr(tf), N, int(n), _lib.ptr(coef),
                _lib.ptr(q), _lib.ptr(dq), _lib.ptr(ddq), _lib.stream_ptr(dev)))
        return coef, q, dq, ddq
