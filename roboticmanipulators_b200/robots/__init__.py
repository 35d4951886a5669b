from .fanuc import Fanuc165F, CANONICAL_8, ALL_12  # noqa: F401
from .rrr_robot import RRRRobot  # noqa: F401
from .rr_robot import RRRobot  # noqa: F401
