"""Quintic polynomial profiles: drop-in for utils/polynomial_generator.py (host NumPy, as in
the reference; listed under "next" in SURVEY.md section 8f, not on the accelerated path)."""
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
