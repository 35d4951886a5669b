"""RRRRobot: drop-in for robots/rrr_robot.py of the reference (chain RzTzRyTxRyTx)."""
import numpy as np
import torch

from ._arm import _SerialArm


class RRRRobot(_SerialArm):
    """3R spatial arm.  Class attributes as in rrr_robot.py:36-39."""

    ls = (1.0, 1.0, 1.0)
    qs_lim_deg = ((-360.0, 360.0),
                  (-360.0, 360.0),
                  (-360.0, 360.0))
    _nq = 3
    _fk_fn, _ik_fn = "rk_rrr_fk_jac_", "rk_rrr_ik_"
    _canonical = ((1, 1), (1, -1), (-1, 1), (-1, -1))

    def __init__(self, T_base=None, T_tool=None):
        self.set_transforms(T_base, T_tool)
        self._generate_value_pairs()
        self._calculate_limits_radians()

    def _lengths(self):
        return self.ls

    def inverse_kinematics(self, T, m=1, k=1):
        """rrr_robot.py:137-201; np.ndarray[3], zeros on failure."""
        if abs(m) != 1:
            print("[WARNING] m can only be -1 or 1. Defaulting to 1")
            m = 1
        if abs(k) != 1:
            print("[WARNING] k can only be -1 or 1. Defaulting to 1")
            k = 1
        return self._ik_single(T, (m, k))

    def move_joints(self, qs):
        """rrr_robot.py:203-224 without the plotting: FK of every row of qs (n x 3) in one
        launch.  Returns the list of 4x4 poses the reference would have drawn."""
        q = torch.tensor(np.asarray(qs, dtype=np.float64).reshape(-1, 3), device="cuda")
        return list(self.forward_kinematics_batch(q, layout="aos").cpu().numpy())

    def move_via_points(self, pts):
        """rrr_robot.py:226-280 without the plotting: IK (m = k = 1) of T_IK = [I | pt] for
        every point, one launch.  Returns np.array(qs).T (3 x n) like the reference."""
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        T = np.tile(np.eye(4), (len(pts), 1, 1))
        T[:, :3, 3] = pts
        q, _ = self.inverse_kinematics_batch(torch.tensor(T, device="cuda"), [(1, 1)], layout="aos")
        return q[:, 0].cpu().numpy().T
