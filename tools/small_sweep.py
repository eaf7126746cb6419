#!/usr/bin/env python3
"""Tuning sweep (GPU): microseconds per step of the small-grid kernels (step_cluster_kernel, step_loop_kernel) on the
reference's shipped cases.  usage: small_sweep.py steps grid[,grid...] kernel[:sync[:arith]] ...
(sync = LBM_CL_SYNC of step_cluster_kernel).  Also checks the final pressure field of each strict run against the
golden SerialCode fixture when steps == the case's maxIters."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    from lbm_asynchronous_b200.inputs import read_params, read_obstacles
    from lbm_asynchronous_b200.lattice import make_param

    steps_arg = int(sys.argv[1])
    gin = os.path.join(ROOT, "tests", "golden", "inputs")
    for grid in sys.argv[2].split(","):
        if os.path.exists(os.path.join(gin, f"input_{grid}.params")):
            p = read_params(os.path.join(gin, f"input_{grid}.params"))
            obst = read_obstacles(os.path.join(gin, f"obstacles_{grid}.dat"), p.nx, p.ny)
        else:  # a synthetic channel of that size
            nx, ny = (int(v) for v in grid.split("x"))
            p = make_param(nx, ny, 1000)
            obst = pkg.channel_obstacles(nx, ny)
        steps = steps_arg if steps_arg > 0 else p.maxIters
        ref = None
        for spec in sys.argv[3:]:
            parts = spec.split(":")
            kernel = int(parts[0])
            sync = parts[1] if len(parts) > 1 and parts[1] else "0"
            arith = parts[2] if len(parts) > 2 else "strict"
            os.environ["LBM_CL_SYNC"] = sync
            try:
                with pkg.Lattice(make_param(p.nx, p.ny, steps, p.reynolds_dim, p.density, p.accel, p.omega), obst, kernel=kernel, arith=arith,
                                 ngpus=int(os.environ.get("LBM_SWEEP_GPUS", "1"))) as lat:
                    lat.run(steps)
                    ms = lat.last_run_ms()
                    pr = lat.pressure()
                    sums = lat.tot_u_sums()[0]
                tot = sums[:, 0] + (sums[:, 1] << 24)
                same = ""
                if arith == "strict":
                    if ref is None:
                        ref = (pr.copy(), tot.copy())
                        same = "(first strict run: the others are compared with it)"
                    else:
                        same = "state %s sums %s" % ("identical" if np.array_equal(pr.view(np.uint32), ref[0].view(np.uint32)) else "DIFFERS",
                                                     "identical" if np.array_equal(tot, ref[1]) else "DIFFER")
                print(f"{grid} gpus={os.environ.get('LBM_SWEEP_GPUS', '1')} steps={steps} kernel={kernel} sync={sync} {arith}: {ms / steps * 1e3:7.3f} us/step "
                      f"{p.nx * p.ny * steps / ms / 1e6:7.2f} GLUPS {same}", flush=True)
            except Exception as ex:
                print(f"{grid} kernel={kernel} sync={sync} {arith}: FAILED {ex}", flush=True)


if __name__ == "__main__":
    main()
