"""Sharding of independent streams / frame batches over GPUs.

The path has no exchange step: frames and streams are independent, so ranks never talk on the data
path.  torch.distributed is used only for the bench's barrier and max-over-ranks timing.
"""
from __future__ import annotations


def streams_for_rank(n_streams: int, rank: int, world: int) -> list[int]:
    """stream s lives on GPU s % world: keeps a stream's frame order on one device/handle."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """contiguous split of one stream's frame batch (used when a single stream is spread out)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per, extra = divmod(n_frames, world)
    start = rank * per + min(rank, extra)
    return range(start, start + per + (1 if rank < extra else 0))
