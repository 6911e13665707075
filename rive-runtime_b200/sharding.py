"""Multi-GPU partitioning of the render path. One process per
GPU (torch.distributed / NCCL over NVLink for the plumbing).

The path shards in exactly two ways:

* animation frames / artboard instances: independent units, frame i -> rank
  i mod N, NO collective on the data path (optionally the finished RGBA8
  frames are gathered to rank 0);
* one very large frame: disjoint horizontal bands of screen tiles; every rank
  receives the same flush inputs, renders only its band (by narrowing
  renderTargetUpdateBounds, which the tile rasteriser honours), and ONE
  gather of W x (H/N) x 4 bytes per rank composites the frame.

Paths of one normal frame do NOT shard (compositing is order dependent per
pixel): replicas only.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import trace as T

TILE = 16


def frames_for_rank(frame_count: int, rank: int, world_size: int) -> List[int]:
    """Frame i is rendered by rank i mod N."""
    return list(range(rank, frame_count, world_size))


def band_for_rank(height: int, rank: int, world_size: int, tile: int = TILE) -> Tuple[int, int]:
    """Rows [r0, r1) of the screen owned by `rank`: whole tile rows, as evenly
    as possible, the last band absorbing the remainder rows."""
    tile_rows = (height + tile - 1) // tile
    t0 = tile_rows * rank // world_size
    t1 = tile_rows * (rank + 1) // world_size
    return min(t0 * tile, height), min(t1 * tile, height)


def restrict_to_band(desc: T.FlushDesc, band: Tuple[int, int]) -> T.FlushDesc:
    """Copy of a flush descriptor whose update bounds are clipped to the band."""
    d = T.FlushDesc.from_buffer_copy(desc)
    d.update_bounds[1] = max(desc.update_bounds[1], band[0])
    d.update_bounds[3] = min(desc.update_bounds[3], band[1])
    if d.update_bounds[3] < d.update_bounds[1]:
        d.update_bounds[3] = d.update_bounds[1]
    return d


def gather_bands(local_rows, height: int, width: int, dst_rank: int = 0, group=None, out_frame=None):
    """Gather each rank's band (a [rows, width, 4] uint8 tensor on the rank's device -- CUDA
    with NCCL over NVLink, CPU with gloo) into the full frame on dst_rank.

    When every band has the same number of rows (the usual case: whole tile rows divide evenly),
    the collective lands each band directly in its rows of the composite -- the receive buffers
    ARE row ranges of `out_frame` (allocated here if not given), no staging copy on either side.
    Otherwise rows are padded to the tallest band and copied out afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bands = [band_for_rank(height, r, world) for r in range(world)]
    rows = [b[1] - b[0] for b in bands]
    frame = None
    if rank == dst_rank:
        frame = out_frame if out_frame is not None else torch.empty((height, width, 4), dtype=torch.uint8, device=local_rows.device)
    if len(set(rows)) == 1:
        dist.gather(local_rows.contiguous(), [frame[r0:r1] for r0, r1 in bands] if rank == dst_rank else None, dst=dst_rank, group=group)
        return frame
    max_rows = max(rows)
    padded = torch.zeros((max_rows, width, 4), dtype=torch.uint8, device=local_rows.device)
    padded[: local_rows.shape[0]] = local_rows
    out = [torch.empty_like(padded) for _ in range(world)] if rank == dst_rank else None
    dist.gather(padded, out, dst=dst_rank, group=group)
    if rank != dst_rank:
        return None
    for r, (r0, r1) in enumerate(bands):
        frame[r0:r1] = out[r][: r1 - r0]
    return frame


def gather_frames(local_frames: Sequence, frame_count: int, dst_rank: int = 0, group=None):
    """Collect round-robin sharded frames (rank r holds frames r, r+N, ...) on
    dst_rank in frame order. Not on the render path; for verification."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered: Optional[list] = [None] * world if rank == dst_rank else None
    payload = [f.cpu().numpy() if hasattr(f, "cpu") else np.asarray(f) for f in local_frames]
    dist.gather_object(payload, gathered, dst=dst_rank, group=group)
    if rank != dst_rank:
        return None
    frames = [None] * frame_count
    for r in range(world):
        for k, i in enumerate(frames_for_rank(frame_count, r, world)):
            frames[i] = gathered[r][k]
    return frames
