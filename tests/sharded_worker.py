"""torchrun worker of test_one_process_per_gpu_under_torchrun: each rank owns one slab on its own GPU;
rank 0 also computes the single-GPU result and compares."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    orc = entry.load_oracle()
    from lbm_asynchronous_b200.lattice import make_param
    from lbm_asynchronous_b200.sharded import ShardedLattice

    gin = os.path.join(ROOT, "tests", "golden", "inputs")
    grid = sys.argv[2] if len(sys.argv) > 2 else "1024x1024"
    kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p = orc.read_params(os.path.join(gin, f"input_{grid}.params"))
    obst = orc.read_obstacles(os.path.join(gin, f"obstacles_{grid}.dat"), p.nx, p.ny)
    iters = 700
    param = make_param(p.nx, p.ny, iters, p.reynolds_dim, p.density, p.accel, p.omega)
    sh = ShardedLattice(param, lambda r0, r1: obst[r0:r1], local, kernel=kernel)
    sh.run(300)
    av1 = sh.av_vels()
    sh.run(400)
    av2 = sh.av_vels()
    cells = sh.cells(dst=0)
    sh.close()
    if rank == 0:
        with pkg.Lattice(param, obst, ngpus=1) as lat:
            lat.run(iters)
            ref_cells, ref_av = lat.cells(), lat.av_vels()
        av = np.concatenate([av1, av2])
        np.savez(sys.argv[1], cells_equal=np.array_equal(cells.view(np.uint32), ref_cells.view(np.uint32)),
                 av_equal=np.array_equal(av.view(np.uint32), ref_av.view(np.uint32)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
