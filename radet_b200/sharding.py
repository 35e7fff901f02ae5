"""Data-parallel plumbing of the hot path: images are sharded over ranks, nothing else is exchanged.

Reference behaviour: RADetHead keeps its normalisers rank-local (radet_head.py:254-259), so reference parity
needs NO collective on the path.  `reduce_mean_` is the opt-in FCOS/ATSS-style normaliser sync
(core/utils/dist_utils.py:63-69, used by atss_head.py:278,296 but not by RADetHead).
"""
import torch
import torch.distributed as dist


def image_range(rank: int, world_size: int, images_per_rank: int, first: int = 0):
    """Images [lo, hi) owned by `rank` (the DDP sampler's contiguous split, datasets/builder.py:112-122)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    lo = first + rank * images_per_rank
    return lo, lo + images_per_rank


def reduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place reduce_mean: t <- sum_over_ranks(t / world).  No-op when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return t
    world = dist.get_world_size(group)
    if world == 1:
        return t
    t /= world
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
