#!/usr/bin/env python3
"""Development probe (GPU): how long lbm_create takes for the benchmark lattice, several times in one process, with the
library's own phase timings (LBM_DEBUG=1).  usage: create_probe.py [nx ny repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LBM_DEBUG"] = "1"
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    from lbm_asynchronous_b200.lattice import make_param

    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    obst = pkg.pack_obstacles(pkg.channel_obstacles(nx, ny))
    keep = None
    for i in range(reps):
        t0 = time.perf_counter()
        lat = pkg.Lattice(make_param(nx, ny, 10), obst, ngpus=1)
        lat.sync()
        t1 = time.perf_counter()
        lat.run(4)
        lat.sync()
        t2 = time.perf_counter()
        print(f"create {i}: {1e3 * (t1 - t0):8.2f} ms, run(4) {1e3 * (t2 - t1):8.2f} ms, another lattice alive: {keep is not None}", flush=True)
        if i == 1:
            keep = lat  # from now on a second lattice of the same size stays alive, as in bench.py's e2e leg
        else:
            lat.close()
    if keep is not None:
        keep.close()


if __name__ == "__main__":
    main()
