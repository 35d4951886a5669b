"""Sharding of independent units (poses, targets, trajectories, measurements) across the
ranks of one node: contiguous blocks, no data-path collective."""


def shard_bounds(n, rank, world):
    """[lo, hi) of rank `rank` when n units are split into `world` contiguous blocks whose
    sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
