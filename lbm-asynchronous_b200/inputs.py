"""Readers of the reference's input formats and its acceptance metric, for Python callers (bench.py's
parity leg, tests).  The C host program has its own parser (host/d2q9-bgk.c); this mirrors it.

  params file   : nx ny maxIters reynolds_dim density accel omega, one per line (SerialCode/d2q9-bgk.c:480-506)
  obstacle file : lines `x y 1` (SerialCode/d2q9-bgk.c:588-601)
  check metric  : worst 100 * (ref - sim) / sim over the series (check/check.py:82-100)
"""
from __future__ import annotations

import numpy as np

from .lattice import make_param


def read_params(path: str):
    with open(path) as fh:
        tok = fh.read().split()
    if len(tok) < 7:
        raise ValueError(f"{path}: expected 7 values (nx ny maxIters reynolds_dim density accel omega)")
    return make_param(int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3]), float(tok[4]), float(tok[5]), float(tok[6]))


def read_obstacles(path: str, nx: int, ny: int) -> np.ndarray:
    """int32[ny, nx], 1 = blocked; same range checks as the reference."""
    out = np.zeros((ny, nx), dtype=np.int32)
    a = np.loadtxt(path, dtype=np.int64, ndmin=2)
    if a.size:
        if a.shape[1] != 3:
            raise ValueError("expected 3 values per line in obstacle file")
        if (a[:, 0] < 0).any() or (a[:, 0] > nx - 1).any():
            raise ValueError("obstacle x-coord out of range")
        if (a[:, 1] < 0).any() or (a[:, 1] > ny - 1).any():
            raise ValueError("obstacle y-coord out of range")
        if (a[:, 2] != 1).any():
            raise ValueError("obstacle blocked value should be 1")
        out[a[:, 1], a[:, 0]] = 1
    return out


def check_metric(ref, sim) -> float:
    """check.py's number: the (signed) percentage difference 100*(ref-sim)/sim of largest magnitude."""
    ref = np.asarray(ref, dtype=np.float64).ravel()
    sim = np.asarray(sim, dtype=np.float64).ravel()
    with np.errstate(divide="ignore", invalid="ignore"):
        pct = 100.0 * (ref - sim) / sim
    if pct.size == 0:
        return 0.0
    if not np.all(np.isfinite(pct)):
        return float("nan")
    return float(pct[np.argmax(np.abs(pct))])
