"""CPU oracle: vectorised float64 NumPy restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the
checker / reported CPU baseline.  The product (roboticmanipulators_b200/) never
imports it and has no CPU fallback.

Parity pinning: every function below is checked in tests/test_oracle_golden.py
against golden vectors produced by running the REAL reference
(tests/golden/make_golden.py, run in the build container against
/root/reference with three non-arithmetic shims) -- "parity pinned".

Each function cites the reference file:line it restates (paths relative to the
reference repo Sarrasor/RoboticManipulators).  The arithmetic is the
reference's: ordinary left-to-right double-precision 4x4 chain products
(utils/robo_math.py:395-397), libm sqrt/acos/atan2/sin/cos, the same thresholds
and the same failure sentinels.  Everything takes a leading batch axis.
"""
import math
import re

import numpy as np

EPS = 1e-5  # robots/robot.py:17

# ---------------------------------------------------------------------------
# A6: elementary transforms and chain products (utils/robo_math.py:363-608)
# ---------------------------------------------------------------------------
_AX = {"x": 0, "y": 1, "z": 2}
VALID_TOKENS = tuple(f"{k}{a}{s}" for k in "TR" for a in "xyz" for s in ("", "i", "d"))


def tokens_from_sequence(sequence):
    """utils/robo_math.py:410-412 (regex split on capital letters)."""
    return [s for s in re.split("([A-Z][^A-Z]*)", sequence) if s]


def elem(token, v):
    """Elementary 4x4 for `token` at value(s) v -> [..., 4, 4].

    utils/robo_math.py:466-608: T* translations, R* rotations, '*i' inverses
    (translation by -v; rotation transposed), '*d' derivative matrices
    (T*d: constant unit generator, R*d: dR/dv with a ZERO [3,3] entry)."""
    if token not in VALID_TOKENS:
        raise ValueError("Unknown token")
    v = np.asarray(v, dtype=np.float64)
    M = np.zeros(v.shape + (4, 4))
    kind, ax, suf = token[0], _AX[token[1]], token[2:]
    if kind == "T":
        if suf == "d":                       # robo_math.py:556-581
            M[..., ax, 3] = 1.0
            return M
        M[..., 0, 0] = M[..., 1, 1] = M[..., 2, 2] = M[..., 3, 3] = 1.0
        M[..., ax, 3] = -v if suf == "i" else v   # :466-503
        return M
    c, s = np.cos(v), np.sin(v)
    i, j = (ax + 1) % 3, (ax + 2) % 3        # rotation acts in the (i, j) plane
    if suf == "d":                           # robo_math.py:583-608
        M[..., i, i] = -s
        M[..., i, j] = -c
        M[..., j, i] = c
        M[..., j, j] = -s
        return M
    M[..., ax, ax] = 1.0
    M[..., 3, 3] = 1.0
    M[..., i, i] = c
    M[..., j, j] = c
    if suf == "i":                           # transpose, :532-542
        M[..., i, j] = s
        M[..., j, i] = -s
    else:                                    # :505-530
        M[..., i, j] = -s
        M[..., j, i] = s
    return M


def chain(sequence, values):
    """Transformation(sequence, values).transformation, batched.

    utils/robo_math.py:364-397: left-to-right product of the token matrices.
    values: [..., ntok]."""
    tokens = tokens_from_sequence(sequence) if isinstance(sequence, str) else list(sequence)
    values = np.asarray(values, dtype=np.float64)
    if values.shape[-1] != len(tokens):
        raise ValueError("Transformation and values sizes do not match")
    T = np.broadcast_to(np.eye(4), values.shape[:-1] + (4, 4)).copy()
    for i, tok in enumerate(tokens):
        T = T @ elem(tok, values[..., i])
    return T


def inverse_sequence(sequence):
    """utils/robo_math.py:418-427 / :139-148: reverse tokens, toggle 'i'."""
    toks = tokens_from_sequence(sequence)
    return "".join(t[:-1] if t.endswith("i") else t + "i" for t in toks[::-1])


def jacobian_column(J):
    """utils/robo_math.py:430-438."""
    return np.stack([J[..., 0, 3], J[..., 1, 3], J[..., 2, 3],
                     J[..., 2, 1], J[..., 0, 2], J[..., 1, 0]], axis=-1)


def skew(x, y, z):
    """utils/robo_math.py:441-446."""
    x, y, z = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float),
                                  np.asarray(z, float))
    S = np.zeros(x.shape + (3, 3))
    S[..., 0, 1], S[..., 0, 2] = -z, y
    S[..., 1, 0], S[..., 1, 2] = z, -x
    S[..., 2, 0], S[..., 2, 1] = -y, x
    return S


# ---------------------------------------------------------------------------
# A7: Jacobians (utils/jacobians.py)
# ---------------------------------------------------------------------------
def jacobian_numerically(sequence, joint_indices, values, T_base=None, T_tool=None):
    """JacobianCalculator.calculate_numerically (utils/jacobians.py:166-184).

    One derivative chain per active token (:94-101); column =
    get_jacobian_column(T_base * dT * T_tool * R^T).  -> [..., 6, n_active]."""
    tokens = tokens_from_sequence(sequence)
    if len(tokens) != len(joint_indices):
        raise ValueError("Size of indices does not match")
    values = np.asarray(values, dtype=np.float64)
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    R = Tb @ chain(tokens, values) @ Tt
    R = R.copy()
    R[..., 0, 3] = R[..., 1, 3] = R[..., 2, 3] = 0.0
    Rt = np.swapaxes(R, -1, -2)
    cols = []
    for i, (tok, active) in enumerate(zip(tokens, joint_indices)):
        if not active:
            continue
        dtoks = tokens[:i] + [tok + "d"] + tokens[i + 1:]
        Jr = Tb @ chain(dtoks, values) @ Tt @ Rt
        cols.append(jacobian_column(Jr))
    return np.stack(cols, axis=-1)


def jacobian_screw(sequence, joint_indices, values, T_base=None, T_tool=None):
    """Numeric form of JacobianCalculator.calculate_scew (utils/jacobians.py:186-240):
    revolute column [u x (O_n - O_k); u], prismatic [u; 0], u = axis column of
    the world frame BEFORE the joint token, O_n includes T_tool."""
    tokens = tokens_from_sequence(sequence)
    values = np.asarray(values, dtype=np.float64)
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    On = (Tb @ chain(tokens, values) @ Tt)[..., :3, 3]
    cols = []
    for i, (tok, active) in enumerate(zip(tokens, joint_indices)):
        if not active:
            continue
        F = Tb @ chain(tokens[:i], values[..., :i])
        u = F[..., :3, _AX[tok[1]]]
        O = F[..., :3, 3]
        if tok[0] == "R":
            cols.append(np.concatenate([np.cross(u, On - O), u], axis=-1))
        else:
            cols.append(np.concatenate([u, np.zeros_like(u)], axis=-1))
    return np.stack(cols, axis=-1)


# ---------------------------------------------------------------------------
# A0/A1/A2: Fanuc165F (robots/fanuc.py)
# ---------------------------------------------------------------------------
FANUC_SEQ = "TzRzTzTxRyTxRyTzTxRxRyRxTx"                       # fanuc.py:122
FANUC_IND = [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0]            # fanuc_jacobians.py:14
FANUC_LIM_DEG = ((-185.0, 185.0), (-60.0, 76.0), (-156.0, 156.0),
                 (-360.0, 360.0), (-125.0, 125.0), (-360.0, 360.0))  # fanuc.py:36-41
CANONICAL_8 = [(m, k, w) for m in (1, -1) for k in (1, -1) for w in (0, 1)]


class FanucConst:
    """Per-robot constants (fanuc.py:43-111, robot.py:39-57)."""

    def __init__(self, T_base=None, T_tool=None, lengths=None, offsets=None,
                 directions=None):
        self.T_base = np.eye(4) if T_base is None else np.array(T_base, dtype=float)
        self.T_tool = np.eye(4) if T_tool is None else np.array(T_tool, dtype=float)
        self.ls = (346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0) \
            if lengths is None else tuple(float(x) for x in lengths)      # :73
        ang = np.pi / 2
        self.offsets = np.array([0.0, -ang, ang, 0.0, 0.0, 0.0]) \
            if offsets is None else np.asarray(offsets, float)            # :79-80
        self.directions = np.array([1, 1, -1, -1, -1, -1], dtype=float) \
            if directions is None else np.asarray(directions, float)      # :86
        self.d = float(np.sqrt(self.ls[4] * self.ls[4] + self.ls[5] * self.ls[5]))  # :98
        self.dq = float(np.arctan2(self.ls[4], self.ls[5]))                          # :99
        self.lim = np.array([(np.deg2rad(a), np.deg2rad(b))
                             for a, b in FANUC_LIM_DEG])                  # :110-111
        self.eps = EPS


def fanuc_values(qm, ls):
    """13 token values for FANUC_SEQ from MODEL angles qm [..., 6]."""
    qm = np.asarray(qm, float)
    o = np.ones(qm.shape[:-1])
    return np.stack([ls[0] * o, qm[..., 0], ls[1] * o, ls[2] * o, qm[..., 1],
                     ls[3] * o, qm[..., 2], ls[4] * o, ls[5] * o,
                     qm[..., 3], qm[..., 4], qm[..., 5], ls[6] * o], axis=-1)


def fanuc_fk(q, c=None):
    """Fanuc165F.forward_kinematics (fanuc.py:151-179): q' = q*dir + off (:163),
    T = T_base * chain * T_tool (:173)."""
    c = c or FanucConst()
    q = np.asarray(q, float)
    qm = q * c.directions + c.offsets
    return c.T_base @ chain(FANUC_SEQ, fanuc_values(qm, c.ls)) @ c.T_tool


def fanuc_jacobian(q, c=None, model_angles=False):
    """6x6 Jacobian w.r.t. MODEL angles q' through calculate_numerically
    (utils/jacobians.py:166-184; driver fanuc_jacobians.py:17-57).
    q are user angles unless model_angles=True."""
    c = c or FanucConst()
    q = np.asarray(q, float)
    qm = q if model_angles else q * c.directions + c.offsets
    return jacobian_numerically(FANUC_SEQ, FANUC_IND, fanuc_values(qm, c.ls),
                                c.T_base, c.T_tool)


def fanuc_ik(T, m=1, k=1, w=0, c=None, details=None):
    """Fanuc165F.inverse_kinematics (fanuc.py:196-302), batched.

    `details` (a dict, optional) receives the quantities the reference compares against its
    thresholds -- psi_cos (:237), A (:253), r_11 (:272), r_13 (:279) and the user angles before
    the limit test (:292-297) -- so that tests can tell how far a target is from every decision,
    and the lengths that condition the atan2 calls (rxy, |(x', z')|, the size of the target).

    Returns (qs [...,6], valid [...]) where invalid rows are the reference's
    all-zero failure sentinel (:239, :300).  Quirks reproduced: k=-1 keeps
    q0 = atan2(y, x); |A|<eps -> q0=0; wrist-singular -> q3=q4=q5=0; NaN q4
    passes the limit test."""
    c = c or FanucConst()
    if abs(m) != 1:
        m = 1                                           # :210-212
    if abs(k) != 1:
        k = 1                                           # :214-216
    if abs(w) != 1 and w != 0:
        w = 1                                           # :218-220
    T = np.asarray(T, float)
    ls, d, dq = c.ls, c.d, c.dq
    Tz_inv = elem("Tz", -ls[0])                         # :223
    Tx_inv = elem("Tx", -ls[6])                         # :224
    T_0 = Tz_inv @ np.linalg.inv(c.T_base) @ T @ np.linalg.inv(c.T_tool) @ Tx_inv  # :225
    x, y, z = T_0[..., 0, 3], T_0[..., 1, 3], T_0[..., 2, 3]
    with np.errstate(invalid="ignore", divide="ignore"):
        x_prime = k * np.sqrt(x**2 + y**2) - ls[2]      # :228
        z_prime = ls[1] - z                             # :229
        num = x_prime**2 + z_prime**2 - d**2 - ls[3]**2  # :231-232
        den = 2.0 * d * ls[3]
        psi_cos = num / den
        unreachable = np.abs(psi_cos) > 1               # :237
        psi = -m * np.arccos(psi_cos)                   # :241
        beta = np.arctan2(d * np.sin(m * psi), ls[3] + d * np.cos(psi))  # :242-243
        gamma = np.arctan2(z_prime, x_prime)            # :244
        q_0 = np.arctan2(y, x)                          # :246
        q_1 = gamma - m * beta                          # :247
        q_2 = psi + dq                                  # :248
        A = d * np.cos(q_1 + psi) + ls[2] + ls[3] * np.cos(q_1)  # :250-251
        q_0 = np.where(np.abs(A) < c.eps, 0.0, q_0)     # :253-255
        # T_012^-1 (fanuc.py:138-140 + robo_math.py:139-148):
        inv_seq = inverse_sequence("RzTzTxRyTxRyRyiTx")
        o = np.ones_like(q_0)
        vals = np.stack([d * o, dq * o, q_2, ls[3] * o, q_1, ls[2] * o,
                         ls[1] * o, q_0], axis=-1)      # reversed variable order
        vals = np.nan_to_num(vals, nan=0.0)             # unreachable rows are discarded below
        T_012_inv = chain(inv_seq, vals)                # :262
        T_rot = elem("Ry", -dq) @ T_012_inv @ T_0       # :263-264
        r_11, r_21, r_31 = T_rot[..., 0, 0], T_rot[..., 1, 0], T_rot[..., 2, 0]
        r_12, r_13 = T_rot[..., 0, 1], T_rot[..., 0, 2]
        regular = np.abs(r_11 - 1.0) > c.eps            # :272
        q_3 = np.arctan2(r_21, -r_31) - w * np.pi       # :273,276
        q_5 = np.arctan2(r_12, r_13) - w * np.pi        # :274,277
        q_4a = np.arctan2(r_13 / np.cos(q_5), r_11)     # :280
        q_4b = np.arctan2(r_12 / np.sin(q_5), r_11)     # :282
        q_4 = np.where(np.abs(r_13) > c.eps, q_4a, q_4b)
        q_3 = np.where(regular, q_3, 0.0)               # :284-287
        q_4 = np.where(regular, q_4, 0.0)
        q_5 = np.where(regular, q_5, 0.0)
        qs = np.stack([q_0, q_1, q_2, q_3, q_4, q_5], axis=-1)
        qs = (qs - c.offsets) * c.directions            # :290
        out = ((qs < c.lim[:, 0]) | (qs > c.lim[:, 1])).any(axis=-1)  # :292-297
    valid = ~unreachable & ~out
    if details is not None:
        details.update(psi_cos=psi_cos, A=A, r_11=r_11, r_13=r_13, qs_user=qs, unreachable=unreachable,
                       rxy=np.sqrt(x**2 + y**2), rg=np.sqrt(x_prime**2 + z_prime**2),
                       reach=np.sqrt(x**2 + y**2 + z**2) + ls[6])
    qs = np.where(valid[..., None], qs, 0.0)
    return qs, valid


def fanuc_ik_branches(T, branches=CANONICAL_8, c=None):
    """All branches -> (q [..., B, 6], valid [..., B]); order = `branches`."""
    qs, vs = zip(*(fanuc_ik(T, m, k, w, c) for (m, k, w) in branches))
    return np.stack(qs, axis=-2), np.stack(vs, axis=-1)


def jac_det(J):
    """Singularity measure det J (fanuc_jacobians.py:56-57)."""
    return np.linalg.det(J)


# ---------------------------------------------------------------------------
# A3/A4/A5: RRR and RR robots
# ---------------------------------------------------------------------------
RRR_SEQ, RRR_IND = "RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0]   # rrr_robot.py:86-87
RR_SEQ, RR_IND = "RzTxRzTx", [1, 0, 1, 0]               # rr_robot.py:92-93
LIM_360 = np.deg2rad(360.0)


def rrr_values(q, ls=(1.0, 1.0, 1.0)):
    q = np.asarray(q, float)
    o = np.ones(q.shape[:-1])
    return np.stack([q[..., 0], ls[0] * o, q[..., 1], ls[1] * o, q[..., 2], ls[2] * o], -1)


def rrr_fk(q, T_base=None, T_tool=None, ls=(1.0, 1.0, 1.0)):
    """RRRRobot.forward_kinematics (rrr_robot.py:96-120)."""
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    return Tb @ chain(RRR_SEQ, rrr_values(q, ls)) @ Tt


def rrr_jacobian(q, T_base=None, T_tool=None, ls=(1.0, 1.0, 1.0)):
    return jacobian_numerically(RRR_SEQ, RRR_IND, rrr_values(q, ls), T_base, T_tool)


def rrr_ik(T, m=1, k=1, T_base=None, T_tool=None, ls=(1.0, 1.0, 1.0)):
    """RRRRobot.inverse_kinematics (rrr_robot.py:137-201)."""
    if abs(m) != 1:
        m = 1
    if abs(k) != 1:
        k = 1
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    T_0 = np.linalg.inv(Tb) @ np.asarray(T, float) @ np.linalg.inv(Tt)   # :159
    x, y, z = T_0[..., 0, 3], T_0[..., 1, 3], T_0[..., 2, 3]
    with np.errstate(invalid="ignore"):
        x_prime = k * np.sqrt(x**2 + y**2)               # :162
        z_prime = ls[0] - z                              # :163
        cc = (x_prime**2 + z_prime**2 - ls[1]**2 - ls[2]**2) / (2.0 * ls[1] * ls[2])
        unreachable = np.abs(cc) > 1                     # :171
        q_2 = m * np.arccos(cc)                          # :175
        beta = np.arctan2(ls[2] * np.sin(m * q_2), ls[1] + ls[2] * np.cos(q_2))  # :177-178
        q_1 = np.arctan2(z_prime, x_prime) - m * beta    # :179
        A = ls[1] * np.cos(q_1) + ls[2] * np.cos(q_1 + q_2)  # :181
        q_0 = np.where(np.abs(A) < EPS, 0.0, np.arctan2(y, x))  # :183-187
        qs = np.stack([q_0, q_1, q_2], axis=-1)
        out = ((qs < -LIM_360) | (qs > LIM_360)).any(axis=-1)
    valid = ~unreachable & ~out
    return np.where(valid[..., None], qs, 0.0), valid


def rr_values(q, ls=(0.8, 0.8)):
    q = np.asarray(q, float)
    o = np.ones(q.shape[:-1])
    return np.stack([q[..., 0], ls[0] * o, q[..., 1], ls[1] * o], -1)


def rr_fk(q, T_base=None, T_tool=None, ls=(0.8, 0.8)):
    """RRRobot.forward_kinematics (rr_robot.py:102-126)."""
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    return Tb @ chain(RR_SEQ, rr_values(q, ls)) @ Tt


def rr_jacobian(q, T_base=None, T_tool=None, ls=(0.8, 0.8)):
    return jacobian_numerically(RR_SEQ, RR_IND, rr_values(q, ls), T_base, T_tool)


def rr_ik(T, m=1, T_base=None, T_tool=None, ls=(0.8, 0.8)):
    """RRRobot.inverse_kinematics (rr_robot.py:147-195)."""
    if abs(m) != 1:
        m = 1
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    T_0 = np.linalg.inv(Tb) @ np.asarray(T, float) @ np.linalg.inv(Tt)
    x, y = T_0[..., 0, 3], T_0[..., 1, 3]
    with np.errstate(invalid="ignore"):
        cc = (x**2 + y**2 - ls[0]**2 - ls[1]**2) / (2.0 * ls[0] * ls[1])  # :167-169
        unreachable = np.abs(cc) > 1
        q_1 = m * np.arccos(cc)                          # :176
        beta = np.arctan2(ls[1] * np.sin(m * q_1), ls[0] + ls[1] * np.cos(q_1))
        q_0 = np.arctan2(y, x) - m * beta                # :179
        qs = np.stack([q_0, q_1], axis=-1)
        out = ((qs < -LIM_360) | (qs > LIM_360)).any(axis=-1)
    valid = ~unreachable & ~out
    return np.where(valid[..., None], qs, 0.0), valid


# ---------------------------------------------------------------------------
# A8: trapezoidal profile (utils/trapezoidal_generator.py)
# ---------------------------------------------------------------------------
def trap_coefficients(q_init, q_final, dq_max, ddq_max, control_freq=0):
    """generate_coefficients (trapezoidal_generator.py:36-79), scalar.
    Returns (t_1, tau)."""
    delta_q = abs(q_final - q_init)
    dq_max_prime = np.sqrt(delta_q * ddq_max)
    if dq_max_prime < dq_max:                 # triangular (:58-61)
        t_1 = dq_max_prime / ddq_max
        tau = t_1
    else:                                     # trapezoidal (:63-66)
        t_1 = dq_max / ddq_max
        tau = delta_q / dq_max
    if control_freq != 0:                     # :71-77
        n = math.ceil(t_1 * control_freq)
        mm = math.ceil(tau * control_freq)
        t_1 = n / control_freq
        tau = mm / control_freq
    return float(t_1), float(tau)


def trap_times(t_1, tau, n):
    """_get_times (trapezoidal_generator.py:208-227): three linspaces whose
    lengths are int(step*t1), int(step*(tau-t1)), int(step*t1)."""
    step = n / (tau + t_1)
    t_acc = np.linspace(0, t_1, int(step * t_1))
    t_no_acc = np.linspace(t_1, tau, int(step * (tau - t_1)))
    t_dec = np.linspace(tau, tau + t_1, int(step * t_1))
    return t_acc, t_no_acc, t_dec


def trap_sample(q_init, v, a, t_1, tau, n):
    """get_q / get_dq / get_ddq / get_t (trapezoidal_generator.py:110-206) for
    current limits (v, a).  Raises like the reference when the acceleration
    segment has no samples (IndexError at :142)."""
    t_acc, t_no_acc, t_dec = trap_times(t_1, tau, n)
    if len(t_acc) == 0:
        raise IndexError("Not enough points. Please, increase n")
    t_flat = t_no_acc - t_acc[-1]                              # :142
    q_acc = q_init + 0.5 * a * t_acc**2                        # :147
    q_no_acc = q_acc[-1] + v * t_flat                          # :148
    start_q = q_no_acc[-1] if len(q_no_acc) != 0 else q_acc[-1]  # :151-154
    q_dec = start_q + v * t_acc - 0.5 * a * t_acc**2           # :156-157
    dq_acc = a * t_acc                                         # :182
    dq = np.hstack([dq_acc, np.ones(len(t_no_acc)) * v, dq_acc[-1] - dq_acc])  # :183-186
    ddq = np.hstack([a * np.ones(len(t_acc)), np.zeros(len(t_no_acc)),
                     -a * np.ones(len(t_dec))])                # :202-206
    return (np.hstack([q_acc, q_no_acc, q_dec]), dq, ddq,
            np.hstack([t_acc, t_no_acc, t_dec]))


# ---------------------------------------------------------------------------
# A9: equalised profiles, p2p / lin trajectories (utils/trajectory_generator.py)
# ---------------------------------------------------------------------------
def equalized_profiles(qs_0, qs_f, n, v_max, a_max, control_freq=0):
    """_generate_equalized_profiles (trajectory_generator.py:118-159) for ONE
    trajectory.  Returns q [L,J], dq [L,J], ddq [L,J], t [L]."""
    coefs = [trap_coefficients(q0, qf, v_max, a_max, control_freq)
             for q0, qf in zip(qs_0, qs_f)]
    t_1_max = max(c[0] for c in coefs)                         # :144
    tau_max = max(c[1] for c in coefs)                         # :145
    qs, dqs, ddqs, ts = [], [], [], None
    for q0, qf in zip(qs_0, qs_f):
        delta_q = qf - q0
        dq_cur = delta_q / tau_max                             # :150
        ddq_cur = delta_q / (tau_max * t_1_max)                # :151
        q, dq, ddq, ts = trap_sample(q0, dq_cur, ddq_cur, t_1_max, tau_max, n)
        qs.append(q), dqs.append(dq), ddqs.append(ddq)
    return np.array(qs).T, np.array(dqs).T, np.array(ddqs).T, ts


def p2p_batch(qs_0, qs_f, n, v_max, a_max, control_freq=0):
    """generate_p2p_trajectory (trajectory_generator.py:71-92) over a batch,
    padded to n: returns q,dq,ddq [N,n,J], t [N,n], length [N] (0 => the
    reference raised)."""
    qs_0, qs_f = np.asarray(qs_0, float), np.asarray(qs_f, float)
    N, J = qs_0.shape
    q = np.zeros((N, n, J))
    dq, ddq, t = np.zeros_like(q), np.zeros_like(q), np.zeros((N, n))
    length = np.zeros(N, dtype=np.int32)
    for i in range(N):
        try:
            a, b, c_, d_ = equalized_profiles(qs_0[i], qs_f[i], n, v_max, a_max, control_freq)
        except (IndexError, ValueError, ZeroDivisionError):
            continue
        L = len(d_)
        length[i] = L
        q[i, :L], dq[i, :L], ddq[i, :L], t[i, :L] = a, b, c_, d_
    return q, dq, ddq, t, length


def rrr_dq_from_dx(q, dx, ls=(1.0, 1.0, 1.0)):
    """get_dq_from_dx (trajectory_generator.py:161-188) with the RRR J_v^-1 of
    rrr_robot_jacobians.py:14-46: dq = inv(J[:3,:3](q)) . dx, batched."""
    J = rrr_jacobian(q, ls=ls)[..., :3, :3]
    with np.errstate(all="ignore"):
        out = np.full(np.shape(dx), np.nan)
        det = np.linalg.det(J)
        ok = np.abs(det) > 0
        out[ok] = np.linalg.solve(J[ok], np.asarray(dx, float)[ok][..., None])[..., 0]
    return out


def lin_rrr_batch(p_0, p_f, n, dx_max, ddx_max, control_freq=0):
    """generate_lin_trajectory (trajectory_generator.py:94-116) followed by
    RRRRobot.move_via_points (rrr_robot.py:226-280: per sample T_IK=[I|p],
    inverse_kinematics(m=k=1)) and get_dq_from_dx.  Padded like p2p_batch.
    Returns x, dx [N,n,3], t [N,n], q, dq [N,n,3], valid [N,n], length [N]."""
    x, dx, _, t, length = p2p_batch(p_0, p_f, n, dx_max, ddx_max, control_freq)
    T = np.broadcast_to(np.eye(4), x.shape[:2] + (4, 4)).copy()
    T[..., :3, 3] = x
    q, valid = rrr_ik(T, 1, 1)
    dq = rrr_dq_from_dx(q, dx)
    pad = np.arange(n)[None, :] >= length[:, None]
    q[pad], dq[pad], valid[pad] = 0.0, 0.0, False
    return x, dx, t, q, dq, valid, length


# ---------------------------------------------------------------------------
# next: quintic polynomial profiles (utils/polynomial_generator.py)
# ---------------------------------------------------------------------------
def poly_coefficients(q_init, q_final, t_f, t_0=0.0):
    """generate_coefficients (polynomial_generator.py:16-36): inv(A).dot([q0,dq0,ddq0,qf,dqf,ddqf])."""
    def sub(t):                                                   # :38-52
        return np.array([[1, t, t**2, t**3, t**4, t**5],
                         [0, 1, 2 * t, 3 * t**2, 4 * t**3, 5 * t**4],
                         [0, 0, 2, 6 * t, 12 * t**2, 20 * t**3]])
    A = np.vstack((sub(t_0), sub(t_f)))
    return np.linalg.inv(A).dot(np.hstack((np.array(q_init), np.array(q_final))))


def poly_sample(coefs, d, from_t, to_t, n):
    """polynomial_from_coefs / dpolynomial_from_coefs (polynomial_generator.py:54-96)."""
    ts = np.linspace(from_t, to_t, n)
    if d >= len(coefs):
        return np.zeros(n)
    poly = 0
    for i, coef in enumerate(coefs[d:], d):
        poly = poly + coef * math.factorial(d) * math.comb(i, d) * ts**(i - d)
    return poly


# ---------------------------------------------------------------------------
# A10: calibration (utils/model_calibration.py)
# ---------------------------------------------------------------------------
class CalibModel:
    """The data get_reduced_model() leaves on the calibrator
    (model_calibration.py:242-250) plus link lengths / offsets / directions."""

    def __init__(self, r_seq, r_var, r_j_ind, r_nom_par, link_lengths,
                 offsets, directions):
        self.r_seq = str(r_seq)
        self.tokens = tokens_from_sequence(self.r_seq)
        self.r_var = [str(v) for v in r_var]
        self.r_j_ind = [int(i) for i in r_j_ind]
        self.r_par_ind = [1 - i for i in self.r_j_ind]
        self.r_nom_par = [str(v) for v in r_nom_par]
        self.link = {str(k): float(v) for k, v in link_lengths}
        self.offsets = np.asarray(offsets, float)
        self.directions = np.asarray(directions, float)
        self.n_params = sum(self.r_par_ind)


def calib_values(model, qs, parameters):
    """_prepare_evaluation (model_calibration.py:555-583), batched over qs [M,k]."""
    qs = np.asarray(qs, float) * model.directions + model.offsets   # :557
    M = qs.shape[0]
    cols, ji, pi = [], 0, 0
    for name, is_joint in zip(model.r_var, model.r_j_ind):
        if is_joint:
            cols.append(qs[:, ji])
            ji += 1
        else:
            nom = model.link.get(model.r_nom_par[pi], 0.0)          # :577-579
            cols.append(np.full(M, nom + parameters[pi]))           # :580
            pi += 1
    return np.stack(cols, axis=-1)


def calib_pose(model, qs, parameters, T_base=None, T_tool=None):
    """get_pose (model_calibration.py:541-548)."""
    Tb = np.eye(4) if T_base is None else np.asarray(T_base, float)
    Tt = np.eye(4) if T_tool is None else np.asarray(T_tool, float)
    return Tb @ chain(model.tokens, calib_values(model, qs, parameters)) @ Tt


def calib_jacobian(model, qs, parameters, T_base, T_tool):
    """get_jacobian (model_calibration.py:550-553): position rows of
    calculate_numerically over the parameter tokens -> [M,3,P]."""
    vals = calib_values(model, qs, parameters)
    return jacobian_numerically(model.r_seq, model.r_par_ind, vals,
                                T_base, T_tool)[..., :3, :]


def calib_normal_equations(model, qs, pts, parameters, T_base, T_tools):
    """The J^T J / J^T r accumulation of estimate_parameters
    (model_calibration.py:413-422).  pts [M, 3n].  Returns H [P,P], g [P]."""
    P = model.n_params
    H, g = np.zeros((P, P)), np.zeros(P)
    pts = np.asarray(pts, float).reshape(len(qs), -1)
    for j, T_tool in enumerate(T_tools):
        T = calib_pose(model, qs, parameters, T_base, T_tool)
        J = calib_jacobian(model, qs, parameters, T_base, T_tool)
        dp = pts[:, 3 * j:3 * j + 3] - T[:, :3, 3]                  # :420
        H += np.einsum("mrp,mrq->pq", J, J)                         # :421
        g += np.einsum("mrp,mr->p", J, dp)                          # :422
    return H, g


def calib_estimate_parameters(model, qs, pts, parameters, T_base, T_tools):
    """estimate_parameters (model_calibration.py:381-454): inv(H).dot(g) (:430)."""
    H, g = calib_normal_equations(model, qs, pts, parameters, T_base, T_tools)
    return np.linalg.inv(H).dot(g), H, g


def calib_tool_base_equations(model, qs, pts, parameters):
    """A^T A / A^T dp of estimate_tool_base (model_calibration.py:296-332)."""
    pts = np.asarray(pts, float).reshape(len(qs), -1)
    n = pts.shape[1] // 3
    dim = 6 + 3 * n
    T = calib_pose(model, qs, parameters)
    R, p = T[:, :3, :3], T[:, :3, 3]
    M = len(qs)
    A = np.zeros((M, 3 * n, dim))
    pst = np.swapaxes(skew(p[:, 0], p[:, 1], p[:, 2]), -1, -2)      # :312
    for i in range(n):
        A[:, 3 * i:3 * i + 3, 0:3] = np.eye(3)
        A[:, 3 * i:3 * i + 3, 3:6] = pst
        A[:, 3 * i:3 * i + 3, 6 + 3 * i:9 + 3 * i] = R
    dp = pts - np.tile(p, (1, n))                                   # :306
    A_sum = np.einsum("mri,mrj->ij", A, A)                          # :331
    Adp = np.einsum("mri,mr->i", A, dp)                             # :332
    return A_sum, Adp


def compose_base_tools(est):
    """model_calibration.py:357-365."""
    T_base = chain("TxTyTzRxRyRz", est[:6])
    T_tools = [chain("TxTyTz", est[i:i + 3]) for i in range(6, len(est), 3)]
    return T_base, T_tools


def calib_estimate_tool_base(model, qs, pts, parameters):
    """estimate_tool_base (model_calibration.py:268-379), first iteration."""
    A_sum, Adp = calib_tool_base_equations(model, qs, pts, parameters)
    est = np.linalg.inv(A_sum).dot(Adp)                             # :347
    T_base, T_tools = compose_base_tools(est)
    return est, T_base, T_tools, A_sum, Adp


def calib_evaluate(model, qs, pts, parameters, T_base, T_tools):
    """evaluate (model_calibration.py:456-539) -> dict of the 8 metrics."""
    pts = np.asarray(pts, float).reshape(len(qs), -1)
    diffs = []
    for j, T_tool in enumerate(T_tools):
        T = calib_pose(model, qs, parameters, T_base, T_tool)
        diffs.append(pts[:, 3 * j:3 * j + 3] - T[:, :3, 3])
    diff = np.concatenate(diffs, axis=0)
    dist = np.linalg.norm(diff, axis=1)
    n = pts.shape[0] * pts.shape[1] // 3                            # :507
    ad = np.abs(diff)
    return {
        "distance_rms": float(np.sqrt((dist * dist).sum() / n)),
        "x_rms": float(np.sqrt((ad[:, 0]**2).sum() / n)),
        "y_rms": float(np.sqrt((ad[:, 1]**2).sum() / n)),
        "z_rms": float(np.sqrt((ad[:, 2]**2).sum() / n)),
        "max_dist": float(dist.max()),
        "max_diff_x": float(ad[:, 0].max()),
        "max_diff_y": float(ad[:, 1].max()),
        "max_diff_z": float(ad[:, 2].max()),
    }
