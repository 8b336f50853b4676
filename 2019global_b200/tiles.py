"""Interleaved image-tile ownership, host-side mirror of csrc/kernels.h TileMap.

32x32 tiles in row-major tile order; global tile t belongs to rank t % world. A rank's
pixels form a compact array: local tile j = global tile j*world + rank, 1024 entries per
tile (row-major inside the tile), edge tiles padded with out-of-frame entries.
"""
import numpy as np

TILE = 32
TILE_PIX = TILE * TILE


def n_tiles(w, h):
    return ((w + TILE - 1) // TILE) * ((h + TILE - 1) // TILE)


def n_local_tiles(w, h, rank, world):
    n = n_tiles(w, h)
    return (n - rank + world - 1) // world if n > rank else 0


def local_pixels(w, h, rank, world):
    """Global pixel index (y*w+x) of every compact entry of `rank`; -1 for padding."""
    tx_n = (w + TILE - 1) // TILE
    nl = n_local_tiles(w, h, rank, world)
    t = np.arange(nl, dtype=np.int64) * world + rank
    ty, tx = t // tx_n, t % tx_n
    ly, lx = np.divmod(np.arange(TILE_PIX, dtype=np.int64), TILE)
    x = tx[:, None] * TILE + lx[None, :]
    y = ty[:, None] * TILE + ly[None, :]
    g = y * w + x
    g[(x >= w) | (y >= h)] = -1
    return g.reshape(-1)


def untile_numpy(frame, tile_values, w, h, rank, world):
    """Reference scatter used by the CPU tests: frame is (h*w, C), tile_values (n_local_pix, C)."""
    g = local_pixels(w, h, rank, world)
    ok = g >= 0
    frame[g[ok]] = tile_values[: g.shape[0]][ok]
    return frame
