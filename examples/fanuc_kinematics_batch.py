#!/usr/bin/env python
"""Batched counterpart of the reference's `fanuc_kinematics.py` (FK -> all IK branches -> FK again):
N random joint vectors inside the limits go through ONE call (`fk_jac_ik_batch`: pose, 6x6 Jacobian,
det J and the 8 canonical (m, k, w) solutions with validity), then the solutions are pushed back
through the forward kinematics.

The inverse kinematics is the reference's, decision for decision (robots/fanuc.py:196-302), including
what it does NOT do: its k = -1 branches evaluate the same triangle with x' = -r_xy - l_2 and are not
solutions of the pose in general, a wrist-singular pose (|r_11 - 1| <= eps) returns q_3 = q_4 = q_5 = 0,
and 'over the back' joint vectors are not among its branches -- so the k = +1 branches close the loop
to rounding and q itself is recovered for ~63 % of uniformly drawn joint vectors.

  python examples/fanuc_kinematics_batch.py [N]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roboticmanipulators_b200.robots import Fanuc165F


def main(n=100000, seed=0, verbose=True):
    robot = Fanuc165F()
    lim = np.array(robot.qs_lim_rad, dtype=np.float64)
    rng = np.random.default_rng(seed)
    q = torch.tensor(rng.uniform(lim[:, 0], lim[:, 1], size=(n, 6)), device="cuda")

    T, J, det, q_ik, valid = robot.fk_jac_ik_batch(q)            # [n,4,4] [n,6,6] [n] [n,8,6] [n,8]

    # the valid k = +1 branches (b = 4 [m = -1] + 2 [k = -1] + w) are solutions: FK(q_ik) == T
    back = robot.forward_kinematics_batch(q_ik.reshape(n * 8, 6)).reshape(n, 8, 4, 4)
    err = (back - T[:, None]).abs().amax(dim=(2, 3))
    wrist_zero = ~q_ik[:, :, 3:].any(dim=2)                       # the reference's singular convention
    closes = valid & ~wrist_zero
    closes[:, [2, 3, 6, 7]] = False
    err = torch.where(closes, err, torch.zeros_like(err))
    # the pose came from q, so one branch returns q itself (up to 2 pi in the continuous joints)
    d = (q_ik - q[:, None]).abs()
    d = torch.minimum(d, (d - 2 * np.pi).abs())
    found = ((d.amax(dim=2) < 1e-6) & valid).any(dim=1)
    out = {"poses": n, "valid_solutions_per_pose": float(valid.sum()) / n,
           "max_roundtrip_error_mm": float(err.max()), "original_q_recovered": float(found.float().mean()),
           "near_singular_poses": int((det.abs() < 1e3).sum())}
    if verbose:
        for k, v in out.items():
            print(f"{k:28s} {v}")
    return out


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100000)
