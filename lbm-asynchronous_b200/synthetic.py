"""Synthetic channel-flow inputs of the benchmark configurations.

Same map as host/gen_channel.c writes as an obstacle file: rows 0 and ny-1 blocked, x periodic,
every cell of rows 1..ny-3 blocked with probability p by SplitMix64 (seed 42, y-major order, one
draw per cell), row ny-2 (the driven row) clear.  `row0`/`row1` select a row slab without
generating the rest (the stream is counter based: draw i uses state seed + (i+1)*GOLDEN).
"""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64_at(seed: int, index: np.ndarray) -> np.ndarray:
    """The (index+1)-th output of SplitMix64 seeded with `seed` (vectorised, wrapping uint64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (index.astype(np.uint64) + np.uint64(1)) * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def channel_obstacles(nx: int, ny: int, p: float = 0.005, seed: int = 42, row0: int = 0, row1: int | None = None) -> np.ndarray:
    """int32[row1-row0, nx] obstacle map (1 = blocked) of rows [row0, row1) of the nx x ny channel."""
    row1 = ny if row1 is None else row1
    out = np.zeros((row1 - row0, nx), dtype=np.int32)
    chunk = max(1, (1 << 24) // nx)  # rows per batch: keeps temporaries around 128 MB
    for a in range(row0, row1, chunk):
        b = min(row1, a + chunk)
        ys = np.arange(a, b, dtype=np.int64)
        inner = (ys >= 1) & (ys <= ny - 3)
        if inner.any():
            yy = ys[inner]
            idx = ((yy - 1)[:, None] * nx + np.arange(nx, dtype=np.int64)[None, :]).astype(np.uint64)
            r = _splitmix64_at(seed, idx)
            blocked = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) < p
            out[np.nonzero(inner)[0] + (a - row0)] = blocked.astype(np.int32)
        for wall in (0, ny - 1):
            if a <= wall < b:
                out[wall - row0] = 1
    return out


def channel_params(nx: int, ny: int, iters: int):
    """(nx, ny, maxIters, reynolds_dim, density, accel, omega) with the physics values of
    dataSet/input_128x128.params."""
    return dict(nx=nx, ny=ny, maxIters=iters, reynolds_dim=10, density=0.1, accel=0.005, omega=1.85)
