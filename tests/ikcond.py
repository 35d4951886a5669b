"""Conditioning of the Pieper IK (robots/fanuc.py:196-302), used by the parity gates.

The IK is exact arithmetic on a rounded pose (lengths of size L ~ 3000 mm, unit roundoff u), but
some of its steps amplify rounding, and the amplifications multiply along the chain of the
solution:
  q0 = atan2(y, x) (:246)                      d q0 ~ u L / rxy          (wrist centre near the base axis)
  psi = -m acos(psi_cos) (:241)                d psi ~ u / sqrt(1 - psi_cos^2)  (elbow stretched / folded)
  gamma = atan2(z', x') (:244)                 d gamma ~ u L / |(x', z')|
    => q1 = gamma - m beta: cond_1 = 1/sin psi + L/|(x', z')|;  q2 = psi + dq: cond_2 = 1/sin psi
  T_rot = Ry(-dq) T_012(q0, q1, q2)^-1 T_0 (:262-264) inherits the errors of q0, q1, q2:
    cond_R = cond_0 + cond_1 + cond_2
  q3 = atan2(r21, -r31), q5 = atan2(r12, r13) (:273-274): operands of length |sin q4|
    => cond_3 = cond_5 = cond_R / |sin q4|     (wrist nearly singular)
  q4 = atan2(r13 / cos q5, r11) (:280): cond_4 = cond_R, and the REFERENCE's own formula divides the
    rounded r13 by cos q5 (the kernels use hypot(r12, r13), which has no such factor -- near
    q5 = +-pi/2 the difference between the two is the reference's noise), hence  cond_4 / |cos q5|.
SURVEY section 8(d) asks for values within 1e-12 max(1, |q|) on valid branches.  A double carries
u = 1.1e-16, so that bound can only hold while cond <= ~1e3; it is asserted AS IT STANDS for every
value whose cond is below COND_OK (FP64: 500 -- e.g. 1 - |psi_cos| > 2e-6 with a regular wrist, or
|sin q4| > 2e-3 with a regular elbow; > 99 % of uniformly sampled branches), and beyond that the
tolerance grows exactly in proportion to cond.  Never statistically: every single value of every
valid branch is checked against its own bound.  FP32 (u = 6e-8, bound 1e-5): COND_OK = 10.
"""
import numpy as np

from oracle import np_oracle as O

COND_OK_F64 = 500.0      # 1e-12 = 18 u cond  at  cond = 500, u = 2^-53
COND_OK_F32 = 10.0       # 1e-5  = 17 u cond  at  cond = 10,  u = 2^-24


def ik_details(T, branches, c):
    """Oracle answers plus the threshold / conditioning quantities of every (target, branch)."""
    qs, vs, det = [], [], []
    with np.errstate(all="ignore"):
        for (m, k, w) in branches:
            d = {}
            q, v = O.fanuc_ik(T, m, k, w, c, details=d)
            qs.append(q), vs.append(v), det.append(d)
    keys = det[0].keys()
    return np.stack(qs, 1), np.stack(vs, 1), {key: np.stack([d[key] for d in det], 1) for key in keys}


def ik_condition(details):
    """cond [N, B, 6] of the module docstring."""
    with np.errstate(all="ignore"):
        c = np.clip(details["psi_cos"], -1.0, 1.0)
        sin_psi = np.sqrt((1.0 - c) * (1.0 + c))
        r11 = np.clip(details["r_11"], -1.0, 1.0)
        s4 = np.sqrt((1.0 - r11) * (1.0 + r11))
        L = details["reach"]
        c0 = np.maximum(1.0, L / details["rxy"])
        c2 = np.maximum(1.0, 1.0 / sin_psi)
        c1 = c2 + L / details["rg"]
        cR = c0 + c1 + c2
        c35 = cR / s4
        c4 = cR / np.minimum(1.0, np.abs(details["r_13"]) / s4)     # |cos q5| = |r13| / |sin q4|
    return np.stack([c0, c1, c2, c35, c4, c35], axis=-1)


def ik_tolerance(q_ref, details, base=1e-12, cond_ok=COND_OK_F64):
    """Per-value tolerance [N, B, 6] and the mask of values held to the plain `base` bound."""
    cond = np.nan_to_num(ik_condition(details), nan=np.inf)
    amp = np.maximum(1.0, cond / cond_ok)
    return base * amp * np.maximum(1.0, np.abs(np.nan_to_num(q_ref))), cond <= cond_ok


def ik_diff(q, q_ref):
    """|q - q_ref| with q0, q3 and q5 compared modulo 2 pi (atan2 cut at +-pi, and w = +-1 shifts)."""
    d = q - q_ref
    d[..., [0, 3, 5]] = (d[..., [0, 3, 5]] + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d)


def assert_ik_close(q, valid, q_ref, v_ref, details, base=1e-12, cond_ok=COND_OK_F64, min_well=0.9):
    """Identical masks, zeros on failure, NaN where the reference has NaN, and EVERY value of every
    valid branch within its conditioned tolerance.  Returns the fraction of values (on valid
    branches) that were held to the plain `base` bound."""
    assert np.array_equal(valid, v_ref), f"{int((valid != v_ref).sum())} mask differences"
    assert np.array_equal(np.isnan(q), np.isnan(q_ref))
    assert not np.any(np.nan_to_num(q[~valid]))
    tol, well = ik_tolerance(q_ref, details, base, cond_ok)
    err = ik_diff(q, q_ref)
    ok = v_ref[..., None] & ~np.isnan(q_ref)
    viol = ok & (err > tol)
    if viol.any():
        i = np.unravel_index(np.argmax(np.where(ok, err / tol, 0.0)), err.shape)
        raise AssertionError(f"{int(viol.sum())} IK values outside their conditioned tolerance; worst at "
                             f"{i}: err {err[i]:.3e} tol {tol[i]:.3e} psi_cos {details['psi_cos'][i[:2]]:.9f} "
                             f"r_11 {details['r_11'][i[:2]]:.9f} rxy {details['rxy'][i[:2]]:.3e}")
    frac = float(well[ok].mean()) if ok.any() else 1.0
    assert frac >= min_well, frac
    return frac


def mask_margin(details, c, margin):
    """True where every threshold decision of a (target, branch) is at least `margin` away from
    flipping: reachability |psi_cos| <= 1 (:237), |A| < eps (:253), |r_11 - 1| > eps (:272),
    |r_13| > eps (:279) and the joint limits (:292-297; radians)."""
    with np.errstate(all="ignore"):
        clear = np.abs(1.0 - np.abs(details["psi_cos"])) > margin
        clear &= np.abs(np.abs(details["A"]) - c.eps) > margin
        reach = ~details["unreachable"]
        wrist = np.abs(np.abs(details["r_11"] - 1.0) - c.eps) > margin
        wrist &= np.abs(np.abs(details["r_13"]) - c.eps) > margin
        qs = details["qs_user"]
        lim = (np.abs(qs - c.lim[:, 0]) > margin) & (np.abs(qs - c.lim[:, 1]) > margin)
        # an unreachable target is decided by psi_cos alone
        clear &= ~reach | (wrist & lim.all(axis=-1))
    return clear
