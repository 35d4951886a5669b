"""Homogeneous-transform algebra: drop-in for utils/robo_math.py of the reference.

`SymbolicTransformation` and `Transformation` keep the reference's constructor
signatures, properties and static getters (utils/robo_math.py:7-608).  They are host-side
objects (SymPy / NumPy) used for constants and symbolic derivations, exactly as in the
reference.  The data-parallel entry point is `Transformation.batch(sequence, values)`:
one CUDA launch of the generic chain interpreter (rk_chain_fk_jac_f64) over N value rows.

All 18 elementary matrices come from one table-driven builder instead of 18 hand-written
getters; the static `get_*` names the reference exposes are generated from it.
"""
import ctypes as C
import math
import re

import numpy as np
import sympy as sp

_AXIS = {"x": 0, "y": 1, "z": 2}
_KINDS = ("T", "R")
_SUFFIXES = ("", "i", "d")
TOKENS = tuple(k + a + s for s in _SUFFIXES for k in _KINDS for a in "xyz")


def tokens_from_sequence(sequence):
    """Split "TzRzTx..." into tokens at capital letters (robo_math.py:88-90, :410-412)."""
    return [s for s in re.split("([A-Z][^A-Z]*)", sequence) if s]


def invert_sequence(tokens):
    """Token list of the inverse chain: reversed, 'i' suffix toggled (robo_math.py:139-148)."""
    return [t[:-1] if t.endswith("i") else t + "i" for t in tokens[::-1]]


def elementary_rows(token, v, sin, cos):
    """4x4 rows of the elementary transform `token` at value v.

    T*: unit matrix with v on the axis row of the last column ('i': -v; 'd': the constant
    generator with a single 1).  R*: rotation about the axis ('i': transposed; 'd':
    derivative w.r.t. the angle, whose [3,3] entry is 0).  Integer entries for
    translations and float entries for rotations, like robo_math.py:204-360."""
    if token not in TOKENS:
        raise ValueError("Unknown token")
    kind, ax, suf = token[0], _AXIS[token[1]], token[2:]
    if kind == "T":
        rows = [[0] * 4 for _ in range(4)]
        if suf == "d":
            rows[ax][3] = 1
            return rows
        for i in range(4):
            rows[i][i] = 1
        rows[ax][3] = -v if suf == "i" else v
        return rows
    i, j = (ax + 1) % 3, (ax + 2) % 3
    rows = [[0.0] * 4 for _ in range(4)]
    s, c = sin(v), cos(v)
    if suf == "d":
        rows[i][i], rows[i][j], rows[j][i], rows[j][j] = -s, -c, c, -s
        return rows
    rows[ax][ax] = 1.0
    rows[3][3] = 1.0
    rows[i][i] = rows[j][j] = c
    rows[i][j], rows[j][i] = (s, -s) if suf == "i" else (-s, s)
    return rows


def elementary_np(token, v):
    """NumPy float64 elementary matrix (host-side constant folding)."""
    return np.array(elementary_rows(token, float(v), math.sin, math.cos), dtype=np.float64)


def chain_np(tokens, values):
    """Left-to-right product of elementary matrices (robo_math.py:395-397), host NumPy."""
    T = np.eye(4)
    for tok, v in zip(tokens, values):
        T = T @ elementary_np(tok, v)
    return T


def _symbol(symbol, f_of_t=False):
    if f_of_t:
        return sp.Function(symbol)(sp.Symbol('t'))
    return sp.Symbol(symbol)


class _Getters:
    """Static get_Tx ... get_Rzd, get_*_inv and helpers shared by both classes."""

    @staticmethod
    def get_jacobian_column(J):
        """[J03, J13, J23, J21, J02, J10] (robo_math.py:169-178, :429-438)."""
        return sp.Matrix([J[0, 3], J[1, 3], J[2, 3], J[2, 1], J[0, 2], J[1, 0]])

    @staticmethod
    def _tokens_from_sequence(sequence):
        return tokens_from_sequence(sequence)

    def valid_token(self, token):
        return token in TOKENS


class SymbolicTransformation(_Getters):
    """Symbolic chain product; `.frames` holds every prefix product."""

    def __init__(self, sequence_string, variables=None, f_of_t=None, simplify=False):
        self._seq = sequence_string
        self._tokens = tokens_from_sequence(self._seq)
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
        if f_of_t is None:
            self._f_of_t = {var: False for var in self._variables}
        else:
            self._f_of_t = dict(zip(self._variables, f_of_t))
        self._tfs = [sp.eye(4)]
        for token, var in zip(self._tokens, self._variables):
            nxt = self._tfs[-1] * self._from_token(token, var)
            self._tfs.append(sp.simplify(nxt) if simplify else nxt)

    def _from_token(self, token, variable_name):
        if token not in TOKENS:
            raise ValueError("Unknown token")
        v = _symbol(variable_name, self._f_of_t.get(variable_name, False))
        return sp.Matrix(elementary_rows(token, v, sp.sin, sp.cos))

    def __getitem__(self, i):
        return self._tfs[-1][i]

    def __mul__(self, other):
        return SymbolicTransformation(self._seq + other._seq, self._variables + other._variables)

    def __rmul__(self, other):
        return SymbolicTransformation(other._seq + self._seq, other._variables + self._variables)

    @property
    def transformation(self):
        return self._tfs[-1]

    @property
    def frames(self):
        return self._tfs[1:]

    @property
    def variables(self):
        return self._variables

    def substitute(self, var_value_pairs):
        pairs = [(sp.symbols(name), value) for name, value in var_value_pairs]
        self._tfs = [tf.subs(pairs) for tf in self._tfs]

    def evaluate(self, values):
        subs = {sp.symbols(name): value for value, name in zip(values, self._variables)}
        return self.transformation.evalf(subs=subs)

    def evaluate_tuples(self, name_value_tuples):
        subs = {sp.symbols(name): value for name, value in name_value_tuples}
        return self.transformation.evalf(subs=subs)

    def inv(self):
        return SymbolicTransformation("".join(invert_sequence(self._tokens)), self._variables[::-1])

    def get_rotation(self):
        R = sp.Matrix(sp.eye(4))
        R[:3, :3] = self.transformation[:3, :3]
        return R

    def print(self):
        print()
        sp.pprint(self.transformation)
        print()

    @staticmethod
    def _get_symbol(symbol, f_of_t=False):
        return _symbol(symbol, f_of_t)

    @staticmethod
    def get_scew(x, y, z, f_of_t=False):
        x, y, z = (_symbol(s, f_of_t) for s in (x, y, z))
        return sp.Matrix([[0, -z, y], [z, 0, -x], [-y, x, 0]])

    @staticmethod
    def get_inertia_matrix(index=0):
        n = {k: sp.symbols(f"I{k}_{index}") for k in ("xx", "yy", "zz", "xy", "yz", "xz")}
        return sp.Matrix([[n["xx"], n["xy"], n["xz"]],
                          [n["xy"], n["yy"], n["yz"]],
                          [n["xz"], n["yz"], n["zz"]]])


class Transformation(_Getters):
    """Numeric chain product of one value vector (host), plus `batch` on the GPU."""

    def __init__(self, sequence_string, values):
        self._seq = sequence_string
        self._tokens = tokens_from_sequence(self._seq)
        if len(values) != len(self._tokens):
            raise ValueError("Transformation and values sizes do not match")
        for token in self._tokens:
            if token not in TOKENS:
                raise ValueError("Unknown token")
        self._values = values
        try:
            vals = [float(v) for v in values]
            T = [np.eye(4)]
            for tok, v in zip(self._tokens, vals):
                T.append(T[-1] @ elementary_np(tok, v))
            self._tfs = [sp.Matrix(t) for t in T]
        except TypeError:  # symbolic values: same product in SymPy
            self._tfs = [sp.eye(4)]
            for tok, v in zip(self._tokens, values):
                self._tfs.append(self._tfs[-1] * sp.Matrix(elementary_rows(tok, v, sp.sin, sp.cos)))

    def _from_token(self, token, value):
        return sp.Matrix(elementary_rows(token, value, sp.sin, sp.cos))

    @property
    def transformation(self):
        return self._tfs[-1]

    def inv(self):
        return Transformation("".join(invert_sequence(self._tokens)), self._values[::-1])

    @staticmethod
    def get_scew(x, y, z):
        return sp.Matrix([[0, -z, y], [z, 0, -x], [-y, x, 0]])

    @staticmethod
    def get_angles(R):
        """Angles (x, y, z) read off a rotation block with the reference's case split
        (robo_math.py:448-464): R[2,0] == 1 is its gimbal case (x := 0); otherwise y comes from the
        first non-zero of R[2,2], R[1,1], R[1,0], and -- as there -- stays unbound when all three
        vanish."""
        r20 = R[2, 0]
        if r20 == 1:
            return 0, np.arcsin(r20), np.arctan2(R[0, 1], R[0, 2])
        about_z = np.arctan2(R[2, 1], R[2, 2])
        about_x = np.arctan2(R[1, 0], R[0, 0])
        cases = ((R[2, 2], lambda: np.arctan2(r20, R[2, 2] / np.cos(about_z))),
                 (R[1, 1], lambda: np.arctan2(-np.cos(about_x) * r20, R[0, 0])),
                 (R[1, 0], lambda: np.arctan2(-np.sin(about_x) * r20, R[1, 0])))
        for pivot, about_y in cases:
            if pivot != 0:
                return about_x, about_y(), about_z
        raise UnboundLocalError("cannot access local variable 'y' where it is not associated with a value")

    @staticmethod
    def get_Rx_from_degrees(q):
        return Transformation.get_Rx(np.deg2rad(q))

    @staticmethod
    def get_Ry_from_degrees(q):
        return Transformation.get_Ry(np.deg2rad(q))

    @staticmethod
    def get_Rz_from_degrees(q):
        return Transformation.get_Rz(np.deg2rad(q))

    # ------------------------------------------------------------------ batched (GPU)
    @staticmethod
    def batch(sequence_string, values, *, T_base=None, T_tool=None, joint_indices=None,
              layout="auto"):
        """Transformation(sequence, values[i]).transformation for every row i of `values`
        (CUDA tensor [N, ntok], float64) in one launch.  With `joint_indices` also returns
        the 6 x n_active Jacobians of JacobianCalculator.calculate_numerically.
        Returns T [N,4,4] (and J [N,6,na])."""
        return chain_batch(tokens_from_sequence(sequence_string), values, T_base=T_base,
                           T_tool=T_tool, active=joint_indices, layout=layout)


def _make_getter(token, symbolic):
    suffix = {"": "", "i": "_inv", "d": "d"}[token[2:]]
    name = f"get_{token[:2]}{suffix}"
    if symbolic:
        def getter(symbol=token[1] if token[0] == "T" else "q", f_of_t=False):
            return sp.Matrix(elementary_rows(token, _symbol(symbol, f_of_t), sp.sin, sp.cos))
    else:
        def getter(value):
            return sp.Matrix(elementary_rows(token, value, sp.sin, sp.cos))
    getter.__name__ = name
    getter.__doc__ = f"Elementary matrix of token '{token}' (utils/robo_math.py)."
    return name, staticmethod(getter)


for _tok in TOKENS:
    _n, _g = _make_getter(_tok, True)
    setattr(SymbolicTransformation, _n, _g)
    _n, _g = _make_getter(_tok, False)
    setattr(Transformation, _n, _g)


class Point3D():
    """robo_math.py:611-634."""

    def __init__(self, x, y, z):
        self._x, self._y, self._z = x, y, z

    @property
    def x(self):
        return self._x

    @property
    def y(self):
        return self._y

    @property
    def z(self):
        return self._z

    def get_homogeneous(self):
        return np.array([[self._x], [self._y], [self._z], [1.0]])


# --------------------------------------------------------------------------------------
# C-ABI glue for the generic chain interpreter
# --------------------------------------------------------------------------------------
_OPCODE = {"T": 0, "R": 1, "Ti": 2, "Ri": 3, "Td": 4, "Rd": 5}


def opcode(token):
    """kind * 3 + axis of include/rk_b200.h: all 18 tokens of utils/robo_math.py:368-387."""
    if token not in TOKENS:
        raise ValueError("Unknown token")
    return _OPCODE[token[0] + token[2:]] * 3 + _AXIS[token[1]]


def make_chain_struct(tokens, active=None, src=None, cst=None, T_base=None, T_tool=None):
    from .. import _lib
    if len(tokens) > _lib.RK_MAX_TOKENS:
        raise ValueError(f"at most {_lib.RK_MAX_TOKENS} tokens are supported")
    ch = _lib.RkChain()
    ch.ntok = len(tokens)
    for i, tok in enumerate(tokens):
        ch.op[i] = opcode(tok)
        ch.active[i] = int(bool(active[i])) if active is not None else 0
        ch.src[i] = i if src is None else int(src[i])
        ch.cst[i] = 0.0 if cst is None else float(cst[i])
    Tb = np.eye(4) if T_base is None else np.array(T_base, dtype=np.float64)
    Tt = np.eye(4) if T_tool is None else np.array(T_tool, dtype=np.float64)
    _lib.fill(ch.T_base, _lib.xf12(Tb))
    _lib.fill(ch.T_tool, _lib.xf12(Tt))
    ch.has_base = 0 if _lib.is_identity(Tb) else 1
    ch.has_tool = 0 if _lib.is_identity(Tt) else 1
    return ch


def chain_batch(tokens, values, *, T_base=None, T_tool=None, active=None, src=None, cst=None,
                layout="auto", want_T=True):
    """Run rk_chain_fk_jac_f64 / _f32 on a CUDA tensor of token values [N, nvals] (float32 values
    run the FP32 kernel, anything else is computed in float64)."""
    import torch
    from .. import _batch, _lib
    from ..robots.robot import launch
    _lib.require_cuda(values, "values")
    if active is not None and len(active) != len(tokens):
        raise ValueError("Size of indices does not match")
    nvals = values.shape[1]
    if src is None and nvals != len(tokens):
        raise ValueError("Transformation and values sizes do not match")
    dt = torch.float32 if values.dtype == torch.float32 else torch.float64
    if layout == "auto":
        layout = _batch.detect_layout(values, nvals)
    vs, ld = _batch.prep_input(values, nvals, layout, dt)
    n, dev = vs.shape[0], vs.device
    ch = make_chain_struct(tokens, active, src, cst, T_base, T_tool)
    na = int(sum(1 for a in active if a)) if active is not None else 0
    T = _batch.alloc_view(n, (4, 4), layout, dt, dev) if want_T else None
    J = _batch.alloc_view(n, (6, na), layout, dt, dev) if na else None
    fn = _lib.load().rk_chain_fk_jac_f64 if dt == torch.float64 else _lib.load().rk_chain_fk_jac_f32
    launch(fn, dev, C.byref(ch), _lib.ptr(vs), nvals, n, ld, _batch.flags_of(layout), _lib.ptr(T),
           _lib.ptr(J))
    if active is None:
        return T
    return T, J


def chain_dq_from_dx(tokens, values, dx, active, *, rows=None, T_base=None, T_tool=None, src=None,
                     cst=None):
    """dq = J[rows, :]^-1 dx for every sample (rk_chain_dq_from_dx_f64): values CUDA [N, nvals],
    dx CUDA [N, na], na = number of active tokens (<= 6).  Returns (dq [N, na], ok [N] bool)."""
    import torch
    from .. import _lib
    from ..robots.robot import launch
    _lib.require_cuda(values, "values")
    _lib.require_cuda(dx, "dx")
    na = int(sum(1 for a in active if a))
    if not 1 <= na <= 6:
        raise ValueError("between 1 and 6 active tokens are needed for a square system")
    if dx.shape[-1] != na:
        raise ValueError("dx must have one column per active token")
    vs = values.to(torch.float64).contiguous()
    d = dx.to(torch.float64).contiguous()
    n, dev = vs.shape[0], vs.device
    ch = make_chain_struct(tokens, active, src, cst, T_base, T_tool)
    rows_c = None
    if rows is not None:
        if len(rows) != na:
            raise ValueError("rows must name one Jacobian row per active token")
        rows_c = (C.c_int32 * na)(*[int(r) for r in rows])
    dq = torch.empty((n, na), dtype=torch.float64, device=dev)
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    launch(_lib.load().rk_chain_dq_from_dx_f64, dev, C.byref(ch), _lib.ptr(vs), vs.shape[1], rows_c,
           _lib.ptr(d), n, n, _lib.RK_AOS, _lib.ptr(dq), _lib.ptr(ok))
    return dq, ok.view(torch.bool)
