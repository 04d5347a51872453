"""Shard arithmetic of the multi-GPU paths (pure Python; no CUDA needed)."""
from typing import Callable, Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n terms: rank r owns [lo, hi)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partial_layout(world: int, limbs: int) -> Tuple[Callable[[int], int], int]:
    """Layout of the all-gathered MSM partials: rank r's XYZZ partial (limbs u64) at offset r * limbs --
    what plk_msm_combine_partials_dev(curve, d_partials, world, ...) expects."""
    return (lambda r: r * limbs), world * limbs
