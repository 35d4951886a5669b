"""Conditioning of the Pieper IK (robots/fanuc.py:196-302), used by the parity gates.

The IK is exact arithmetic on a rounded pose, but two steps amplify rounding:
  * psi = -m acos(psi_cos) (:241): d psi = d psi_cos / sqrt(1 - psi_cos^2) -- the elbow nearly
    stretched or folded; psi feeds q1, q2 and, through T_rot (:262-264), the whole wrist;
  * q3 = atan2(r21, -r31), q5 = atan2(r12, r13) (:273-274): their operands have length |sin q4|,
    so d q3, d q5 = d r / |sin q4| -- the wrist nearly singular.
SURVEY section 8(d) asks for values within 1e-12 max(1, |q|) on valid branches.  That bound is
asserted as it stands wherever  1 - |psi_cos| > 1e-6  and  |sin q4| > 1e-3; closer to either
singularity the tolerance grows exactly with the amplification factor (never statistically:
every single value is checked against its own bound).
"""
import numpy as np

from oracle import np_oracle as O

ELBOW_OK = np.sqrt(1.0 - (1.0 - 1e-6) ** 2)   # sqrt(1 - c^2) at 1 - |c| = 1e-6  (1.414e-3)
WRIST_OK = 1e-3


def ik_details(T, branches, c):
    """Oracle answers plus the threshold quantities of every (target, branch)."""
    qs, vs, det = [], [], []
    with np.errstate(all="ignore"):
        for (m, k, w) in branches:
            d = {}
            q, v = O.fanuc_ik(T, m, k, w, c, details=d)
            qs.append(q), vs.append(v), det.append(d)
    keys = det[0].keys()
    return np.stack(qs, 1), np.stack(vs, 1), {key: np.stack([d[key] for d in det], 1) for key in keys}


def ik_tolerance(q_ref, details, base=1e-12, elbow_ok=ELBOW_OK, wrist_ok=WRIST_OK):
    """Per-value tolerance [N, B, 6] (see the module docstring)."""
    with np.errstate(all="ignore"):
        c = np.clip(details["psi_cos"], -1.0, 1.0)
        sin_psi = np.sqrt((1.0 - c) * (1.0 + c))
        s4 = np.sqrt(np.maximum(0.0, (1.0 - details["r_11"]) * (1.0 + details["r_11"])))
        a_e = np.maximum(1.0, elbow_ok / sin_psi)
        a_w = np.maximum(1.0, wrist_ok / s4)
    amp = np.stack([np.ones_like(a_e), a_e, a_e, a_e * a_w, a_e, a_e * a_w], axis=-1)
    return base * amp * np.maximum(1.0, np.abs(np.nan_to_num(q_ref))), (a_e == 1.0) & (a_w == 1.0)


def ik_diff(q, q_ref):
    """|q - q_ref| with q3 and q5 compared modulo 2 pi (atan2 cut at +-pi, and w = +-1 shifts)."""
    d = q - q_ref
    d[..., [3, 5]] = (d[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d)


def assert_ik_close(q, valid, q_ref, v_ref, details, base=1e-12, elbow_ok=ELBOW_OK, wrist_ok=WRIST_OK,
                    min_well=0.9):
    """Identical masks, zeros on failure, NaN where the reference has NaN, and EVERY value of every
    valid branch within its conditioned tolerance.  Returns the fraction of well-conditioned
    branches (those held to the plain `base` bound)."""
    assert np.array_equal(valid, v_ref), f"{int((valid != v_ref).sum())} mask differences"
    assert np.array_equal(np.isnan(q), np.isnan(q_ref))
    assert not np.any(np.nan_to_num(q[~valid]))
    tol, well = ik_tolerance(q_ref, details, base, elbow_ok, wrist_ok)
    err = ik_diff(q, q_ref)
    ok = v_ref[..., None] & ~np.isnan(q_ref)
    viol = ok & (err > tol)
    if viol.any():
        i = np.unravel_index(np.argmax(np.where(ok, err / tol, 0.0)), err.shape)
        raise AssertionError(f"{int(viol.sum())} IK values outside their conditioned tolerance; worst at "
                             f"{i}: err {err[i]:.3e} tol {tol[i]:.3e} psi_cos {details['psi_cos'][i[:2]]:.9f} "
                             f"r_11 {details['r_11'][i[:2]]:.9f}")
    frac = float(well[v_ref].mean()) if v_ref.any() else 1.0
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
