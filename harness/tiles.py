"""Host-side mirror of the sort-first tile layout (csrc/fixed.cu k_tiles_pack/_unpack):
tile t (32x32 px, row-major over the screen) is owned by rank t % world; the k-th owned tile of a rank
(t = rank + k*world) occupies 4096 bytes at slot k of that rank's send buffer; the all-gather result is
rank-major."""
from __future__ import annotations

import numpy as np

TILE = 32


def tiles_xy(width: int, height: int):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def tiles_per_rank(width: int, height: int, world: int) -> int:
    tx, ty = tiles_xy(width, height)
    return (tx * ty + world - 1) // world


def owned_mask(width: int, height: int, rank: int, world: int) -> np.ndarray:
    """boolean [H, W]: pixels whose tile belongs to `rank`"""
    tx, _ = tiles_xy(width, height)
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // TILE) * tx + xs // TILE) % world == rank


def pack(color: np.ndarray, rank: int, world: int) -> np.ndarray:
    h, w = color.shape[:2]
    tx, ty = tiles_xy(w, h)
    slots = tiles_per_rank(w, h, world)
    out = np.zeros((slots, TILE, TILE, 4), dtype=np.uint8)
    for k in range(slots):
        t = rank + k * world
        if t >= tx * ty:
            break
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        blk = color[y0:y0 + TILE, x0:x0 + TILE]
        out[k, :blk.shape[0], :blk.shape[1]] = blk
    return out.reshape(-1)


def unpack(gathered: np.ndarray, width: int, height: int, world: int) -> np.ndarray:
    tx, ty = tiles_xy(width, height)
    slots = tiles_per_rank(width, height, world)
    g = gathered.reshape(world, slots, TILE, TILE, 4)
    color = np.zeros((height, width, 4), dtype=np.uint8)
    for t in range(tx * ty):
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        blk = g[t % world, t // world]
        color[y0:y0 + TILE, x0:x0 + TILE] = blk[:min(TILE, height - y0), :min(TILE, width - x0)]
    return color


# ---- sharded host<->device traffic of the N-rank end-to-end path (bench.py) --------------------------------
def upload_shard(nbytes: int, world: int, min_shard: int = 1 << 16):
    """How an input of `nbytes` is split over `world` PCIe links: (slice_bytes, tail_bytes). Rank r uploads
    bytes [r*slice, (r+1)*slice) and the slices are all-gathered; every rank uploads the tail
    [world*slice, nbytes) itself. slice = 0 means "too small to shard": every rank uploads everything."""
    sl = (nbytes // world) & ~0xff
    if sl < min_shard:
        return 0, nbytes
    return sl, nbytes - sl * world


def row_band(height: int, rank: int, world: int):
    """rows [lo, hi) of the assembled image that rank downloads into the shared host buffer"""
    rows = (height + world - 1) // world
    return min(rank * rows, height), min((rank + 1) * rows, height)
