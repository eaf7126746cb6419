#!/usr/bin/env python3
"""Writes the reference results that check.py compares against as `.dat` text files (the reference's
av_vels.dat / final_state.dat formats, SerialCode/d2q9-bgk.c:679-738) from tests/golden/<grid>.npz.

    python tests/ref_check/write_goldens.py OUTDIR [grid ...]

TEST INFRASTRUCTURE.  The fixtures hold the reference's shipped check/*.dat goldens (float64) and the
outputs of the reference's own binaries run on the shipped inputs (oracle/make_fixtures.py); the text
written here has the columns check.py reads (ii, jj, pressure; `tt:` value) exactly as the reference
prints them.  The u_x / u_y / u columns are filled from the SerialCode run where the fixture has them,
else with zeros -- check.py never reads them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GOLD = os.path.join(ROOT, "tests", "golden")
GRIDS = ["128x128", "128x256", "256x256", "1024x1024"]


def read_obstacle_map(grid: str, nx: int, ny: int) -> np.ndarray:
    ob = np.zeros((ny, nx), dtype=np.int64)
    with open(os.path.join(GOLD, "inputs", f"obstacles_{grid}.dat")) as fh:
        for line in fh:
            parts = line.split()
            if len(parts) == 3:
                ob[int(parts[1]), int(parts[0])] = 1
    return ob


def write_grid(outdir: str, grid: str) -> tuple[str, str]:
    fx = np.load(os.path.join(GOLD, f"{grid}.npz"))
    nx, ny = (int(v) for v in grid.split("x"))
    av_path = os.path.join(outdir, f"{grid}.av_vels.dat")
    fs_path = os.path.join(outdir, f"{grid}.final_state.dat")
    av = fx["golden_av_vels"]
    with open(av_path, "w") as fh:
        fh.write("".join(f"{tt}:\t{v:.12E}\n" for tt, v in enumerate(av)))
    if "golden_pressure" in fx.files:
        pressure = np.asarray(fx["golden_pressure"], dtype=np.float64).reshape(ny, nx)
    else:  # no shipped final_state golden for this grid: the reference's own binary, run on the shipped inputs
        pressure = np.asarray(fx["serial_pressure"], dtype=np.float64).reshape(ny, nx)
    planes = []
    for name in ("serial_ux", "serial_uy", "serial_u"):
        planes.append(np.asarray(fx[name], dtype=np.float64).reshape(ny, nx) if name in fx.files else np.zeros((ny, nx)))
    ob = read_obstacle_map(grid, nx, ny)
    with open(fs_path, "w") as fh:
        for jj in range(ny):
            fh.write("".join(
                f"{ii} {jj} {planes[0][jj, ii]:.12E} {planes[1][jj, ii]:.12E} {planes[2][jj, ii]:.12E} {pressure[jj, ii]:.12E} {ob[jj, ii]}\n"
                for ii in range(nx)))
    return av_path, fs_path


def main() -> int:
    if len(sys.argv) < 2:
        print(__doc__, file=sys.stderr)
        return 2
    outdir = sys.argv[1]
    os.makedirs(outdir, exist_ok=True)
    for g in (sys.argv[2:] or GRIDS):
        a, f = write_grid(outdir, g)
        print(f"{g}: {a} {f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
