#!/usr/bin/env python3
"""Tuning sweep (GPU): MLUPS of step kernels variants on one B200.  usage: f2_sweep.py [nx ny steps] [kernel[:seg[:arith]] ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    from lbm_asynchronous_b200.lattice import make_param

    nx, ny, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    obst = pkg.pack_obstacles(pkg.channel_obstacles(nx, ny))
    for spec in sys.argv[4:]:
        parts = spec.split(":")
        kernel = int(parts[0])
        seg = parts[1] if len(parts) > 1 and parts[1] else None
        arith = parts[2] if len(parts) > 2 else "strict"
        if seg:
            os.environ["LBM_F2_SEG"] = seg
        else:
            os.environ.pop("LBM_F2_SEG", None)
        try:
            with pkg.Lattice(make_param(nx, ny, steps), obst, kernel=kernel, arith=arith) as lat:
                lat.run(8)
                lat.sync()
                best = 1e30
                for _ in range(3):
                    lat.run(steps)
                    best = min(best, lat.last_run_ms())
                av = lat.av_vels()
            print(f"{nx}x{ny} kernel={kernel} seg={seg} arith={arith}: {nx * ny * steps / best / 1e6:8.2f} GLUPS  {best / steps * 1e3:9.2f} us/step  av[-1]={av[-1]:.6e}", flush=True)
        except Exception as ex:
            print(f"{nx}x{ny} kernel={kernel} seg={seg} arith={arith}: FAILED {ex}", flush=True)


if __name__ == "__main__":
    main()
