"""Work partitioning of one frame over the GPUs of a node.

The scene is replicated; ranks never exchange anything while tracing.  Two partitions:

  * image tiles   — rank g renders the row blocks b with b % world == g of every sample.  Pixels
                    are disjoint, the RNG uses global pixel coordinates, so the reduced image is
                    BIT-IDENTICAL to a single-GPU render (sum with zeros elsewhere).
  * sample slices — rank g renders samples [first + g*S/world, first + (g+1)*S/world) of the whole
                    frame; the reduced image differs from a single-GPU render only by the order of
                    fp32 additions.

The only collective is one sum-reduce of the float4 accumulation buffer at the end of a render.
"""
from __future__ import annotations

import numpy as np

from . import scene as sc


def row_block_tiles(width: int, height: int, rank: int, world: int, block_rows: int = 8) -> np.ndarray:
    """Interleaved horizontal strips of `block_rows` rows: strip b belongs to rank b % world."""
    tiles = [(0, y, width, min(y + block_rows, height)) for b, y in enumerate(range(0, height, block_rows)) if b % world == rank]
    return np.array(tiles, sc.TILE).reshape(-1)


def sample_slice(first_sample: int, sample_count: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous sample range (first, count) of a rank; ranges tile [first, first + count)."""
    lo = sample_count * rank // world
    hi = sample_count * (rank + 1) // world
    return first_sample + lo, hi - lo


def reduce_accumulation(accum, dst: int = 0):
    """Sum-reduces the float accumulation tensor (torch, on the backend's device) onto `dst` and
    restores alpha = 1 there (every rank's buffer carries alpha 1 on its own pixels)."""
    import torch.distributed as dist

    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() == dst:
        accum.view(-1, 4)[:, 3] = 1.0
    return accum
