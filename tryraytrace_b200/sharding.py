"""Sample-index sharding of a progressive pass across GPUs.

The reference accumulates frame seeds 1, 2, 3, ... on one device (reference src/main.cpp:149,
181, 222).  With G ranks, rank r renders seeds first+r, first+r+G, ... so the union over ranks
is exactly the single-GPU seed set; the per-rank buffers are summed with one all-reduce.
"""
from __future__ import annotations


def frames_for_rank(first_frame_seed: int, n_frames_total: int, rank: int, world: int):
    """(first, count, stride) for `rank`: which frame seeds of the pass it renders."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_frames_total < 0:
        raise ValueError("negative frame count")
    count = (n_frames_total - rank + world - 1) // world if n_frames_total > rank else 0
    return first_frame_seed + rank, count, world


def render_pass_sharded(ctx, accum, width, height, first_frame_seed, n_frames_total, cam, opts=None, dist=None):
    """One progressive pass split over the ranks of `dist` (a torch.distributed module with an
    initialised NCCL group, or None for one GPU): render this rank's frames into `accum`
    (a zeroed CUDA tensor of w*h*4 floats), then all-reduce it once."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    first, count, stride = frames_for_rank(first_frame_seed, n_frames_total, rank, world)
    if count:
        ctx.render(accum, width, height, first, count, cam, opts, frame_stride=stride)
    if dist is not None:
        dist.all_reduce(accum)
    return accum
