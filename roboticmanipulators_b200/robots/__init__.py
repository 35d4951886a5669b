from .fanuc import Fanuc165F, CANONICAL_8, ALL_12  # noqa: F401
