"""Host logic of the one-process-per-GPU path under torch.distributed with the gloo backend,
world_size 2 (no GPU needed): handle exchange around the ring, integer |u| sums, slab gather."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import __graft_entry__ as entry

    entry.load_package()
    from lbm_asynchronous_b200 import capi, sharded

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        # 1. halo handles travel to the ring neighbours
        mine = bytes([rank + 1]) * capi.HALO_HANDLE_BYTES
        south, north = sharded.exchange_handles(mine)
        out["south"], out["north"] = south[0], north[0]
        # 2. integer sums add exactly, whatever the split
        ny, nx, iters = 12, 8, 5
        rng = np.random.default_rng(0)
        v = rng.integers(0, 1 << 38, size=(iters, ny * nx), dtype=np.int64)  # per-cell fixed-point |u|
        starts = capi.partition(ny, world)
        sl = slice(starts[rank] * nx, starts[rank + 1] * nx)
        sums = np.stack([(v[:, sl] & 0xFFFFFF).sum(1), (v[:, sl] >> 24).sum(1)], axis=1).astype(np.int64)
        av = sharded.combine_sums(sums, np.zeros(iters, np.int64), fluid_cells=(starts[rank + 1] - starts[rank]) * nx)
        out["av"] = av
        want = [sharded.av_from_sums(int((v[t] & 0xFFFFFF).sum()), int((v[t] >> 24).sum()), 0, ny * nx) for t in range(iters)]
        out["want"] = np.array(want, dtype=np.float32)
        # 3. slabs gather in rank order
        full = np.arange(ny * nx * 3, dtype=np.float32).reshape(ny, nx, 3)
        got = sharded.gather_rows(full[starts[rank]:starts[rank + 1]], starts, dst=0)
        out["gather_ok"] = (got is None) if rank != 0 else bool(np.array_equal(got, full))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_plumbing_under_gloo(built, world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        o = res[r]
        assert o["south"] == (r - 1) % world + 1 and o["north"] == (r + 1) % world + 1
        assert np.array_equal(o["av"].view(np.uint32), o["want"].view(np.uint32))
        assert o["gather_ok"]
    assert np.array_equal(res[0]["av"].view(np.uint32), res[1]["av"].view(np.uint32))


def test_ring_neighbours(pkg):
    from lbm_asynchronous_b200 import sharded

    assert sharded.ring_neighbours(0, 4) == (3, 1)
    assert sharded.ring_neighbours(3, 4) == (2, 0)
    assert sharded.ring_neighbours(0, 1) == (0, 0)
