"""JacobianCalculator: drop-in for utils/jacobians.py of the reference.

The symbolic methods (`calculate_numeric`, `calculate_scew`) stay SymPy on the host, as in
the reference.  `calculate_numerically(values)` (utils/jacobians.py:166-184) runs a
1-element batch of the generic chain kernel; `calculate_numerically_batch(values[N,ntok])`
is the data-parallel entry point.
"""
import numpy as np
import sympy as sp

from .robo_math import SymbolicTransformation as st
from .robo_math import TOKENS, chain_batch, tokens_from_sequence


class JacobianCalculator():
    """Jacobians of a token chain by matrix differentiation or by screw theory.

    Attributes: simplify, T_base, T_tool (utils/jacobians.py:13-19)."""

    def __init__(self, sequence_string, joint_indices, f_of_t=True, variables=None,
                 T_base=None, T_tool=None, simplify=False):
        self._seq = sequence_string
        self._tokens = tokens_from_sequence(self._seq)
        self._ind = joint_indices
        self.simplify = simplify
        self.set_transforms(T_base, T_tool)
        if len(self._tokens) != len(self._ind):
            raise ValueError("Size of indices does not match")
        if type(f_of_t) == bool:
            self._f_of_t = self._ind if f_of_t else [False] * len(self._ind)
        else:
            self._f_of_t = f_of_t
        if variables is None:
            self._variables, ti, ri = [], 0, 0
            for token in self._tokens:
                if token not in TOKENS:
                    raise ValueError("Unknown transformation")
                if token.startswith('T'):
                    self._variables.append(f"d_{ti}")
                    ti += 1
                else:
                    self._variables.append(f"q_{ri}")
                    ri += 1
        else:
            if len(variables) != len(self._tokens):
                raise ValueError("Transformation and var sizes do not match")
            self._variables = variables
        # one derivative sequence per active token and the prefix before it
        # (utils/jacobians.py:94-105)
        self._seq_numeric, self._seq_skew, self._vars_skew = [], [], []
        for i, (token, active) in enumerate(zip(self._tokens, self._ind)):
            if active:
                self._seq_numeric.append(
                    ''.join(self._tokens[:i] + [f"{token}d"] + self._tokens[i + 1:]))
                self._seq_skew.append(''.join(self._tokens[:i]))
        self._seq_skew.append(''.join(self._tokens))
        self._seq_numeric.append(''.join(self._tokens))
        for seq in self._seq_skew:
            self._vars_skew.append(self._variables[:len(tokens_from_sequence(seq))])

    def set_transforms(self, T_base=None, T_tool=None):
        self.T_base = sp.eye(4) if T_base is None else sp.Matrix(T_base)
        self.T_tool = sp.eye(4) if T_tool is None else sp.Matrix(T_tool)

    # ---------------------------------------------------------------- symbolic (host)
    def calculate_numeric(self, T_base=None, T_tool=None):
        """Symbolic Jacobian by matrix differentiation (utils/jacobians.py:127-164):
        column_i = get_jacobian_column(T_base dT_i T_tool R^T)."""
        T_base = self.T_base if T_base is None else T_base
        T_tool = self.T_tool if T_tool is None else T_tool
        R = T_base * st(self._seq_numeric[-1], self._variables,
                        f_of_t=self._f_of_t).transformation * T_tool
        R[0, 3] = R[1, 3] = R[2, 3] = 0.0
        Rt = R.T
        cols = []
        for seq in self._seq_numeric[:-1]:
            dT = st(seq, self._variables, f_of_t=self._f_of_t).transformation
            cols.append(st.get_jacobian_column(T_base * dT * T_tool * Rt).T)
        res = sp.Matrix(cols).T
        return sp.simplify(res) if self.simplify else res

    def calculate_scew(self, T_base=None, T_tool=None):
        """Symbolic Jacobian by screw theory (utils/jacobians.py:186-240): revolute
        [u x (O_n - O_k); u], prismatic [u; 0]."""
        T_base = self.T_base if T_base is None else T_base
        T_tool = self.T_tool if T_tool is None else T_tool
        full = st(self._seq_skew[-1], self._vars_skew[-1], f_of_t=self._f_of_t).transformation
        O_n = (T_base * full * T_tool)[:3, 3]
        cols = []
        for k in range(len(self._seq_skew) - 1):
            before = T_base * st(self._seq_skew[k], self._vars_skew[k],
                                 f_of_t=self._f_of_t).transformation
            token = tokens_from_sequence(self._seq_skew[k + 1])[len(self._vars_skew[k])]
            u = before[:3, ord(token[1]) - ord('x')]
            if token[0] == 'R':
                cols.append(sp.Matrix([u.cross(O_n - before[:3, 3]), u]).T)
            else:
                cols.append(sp.Matrix([u[0], u[1], u[2], 0, 0, 0]).T)
        res = sp.Matrix(cols).T
        return sp.simplify(res) if self.simplify else res

    # ------------------------------------------------------------------- numeric (GPU)
    def calculate_numerically_batch(self, values, T_base=None, T_tool=None, *, layout="auto",
                                    with_pose=False):
        """Jacobians for N value rows: values CUDA tensor [N, ntok] -> J [N, 6, n_active]
        (and T [N,4,4] with with_pose=True)."""
        T_base = self.T_base if T_base is None else T_base
        T_tool = self.T_tool if T_tool is None else T_tool
        for token, active in zip(self._tokens, self._ind):
            if active and token.endswith("i"):
                raise ValueError("Unknown token")   # the reference has no '<inverse>d' token
        T, J = chain_batch(self._tokens, values,
                           T_base=np.array(T_base, dtype=np.float64),
                           T_tool=np.array(T_tool, dtype=np.float64),
                           active=[int(bool(a)) for a in self._ind], layout=layout)
        return (T, J) if with_pose else J

    def dq_from_dx_batch(self, values, dx, rows=None, T_base=None, T_tool=None):
        """Joint rates from Cartesian rates for N samples: solves J[rows, :] dq = dx in the kernel
        (TrajectoryGenerator.get_dq_from_dx for any chain, utils/trajectory_generator.py:161-188).
        values CUDA [N, ntok], dx CUDA [N, n_active]; rows = the Jacobian rows that form the square
        system (default 0 .. n_active-1).  Returns (dq [N, n_active], ok [N] bool)."""
        from .robo_math import chain_dq_from_dx
        T_base = self.T_base if T_base is None else T_base
        T_tool = self.T_tool if T_tool is None else T_tool
        return chain_dq_from_dx(self._tokens, values, dx, [int(bool(a)) for a in self._ind], rows=rows,
                                T_base=np.array(T_base, dtype=np.float64),
                                T_tool=np.array(T_tool, dtype=np.float64))

    def calculate_numerically(self, values, T_base=None, T_tool=None):
        """utils/jacobians.py:166-184; returns a 6 x n_active sp.Matrix."""
        import torch
        if len(values) != len(self._tokens):
            raise ValueError("Transformation and values sizes do not match")
        v = torch.tensor(np.asarray(values, dtype=np.float64).reshape(1, -1), device="cuda")
        J = self.calculate_numerically_batch(v, T_base, T_tool, layout="aos")
        return sp.Matrix(J[0].cpu().numpy())
