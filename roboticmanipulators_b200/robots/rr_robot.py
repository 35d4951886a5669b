"""RRRobot: drop-in for robots/rr_robot.py of the reference (planar chain RzTxRzTx)."""
from ._arm import _SerialArm


class RRRobot(_SerialArm):
    """2R planar arm (rr_robot.py:16-65)."""

    qs_lim_deg = ((-360.0, 360.0), (-360.0, 360.0))
    _nq = 2
    _fk_fn, _ik_fn = "rk_rr_fk_jac_", "rk_rr_ik_"
    _canonical = (1, -1)

    def __init__(self, T_base=None, T_tool=None, lengths=None, save=True):
        """`save` is accepted for compatibility; there is no pickle cache."""
        self.set_transforms(T_base, T_tool)
        self.set_lengths(lengths)
        self._generate_value_pairs()
        self._calculate_limits_radians()
        self._save = save

    def _lengths(self):
        return self._ls

    def inverse_kinematics(self, T, m=1):
        """rr_robot.py:147-195; np.ndarray[2], zeros on failure."""
        if abs(m) != 1:
            print("[WARNING] m can only be -1 or 1. Defaulting to 1")
            m = 1
        return self._ik_single(T, m)
