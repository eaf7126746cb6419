"""One process per GPU: the grid's row slabs over the ranks of a `torch.distributed` group.

Replaces the reference's MPI set-up (rank/size, row decomposition, obstacle scatter, final gather and
MPI_Reduce; MPI/d2q9-bgk.c:135-151, 661-688, 730-829, 265-309).  torch.distributed is plumbing only:
it carries the 128-byte halo handles at start-up and the integer |u| sums / final rows at the end.
The per-step halo exchange is NOT a collective: each rank's step kernel stores its boundary rows
straight into the neighbour GPU's halo ring over NVLink (lbm_kernels.cuh).

The helpers that hold the host logic (`ring_neighbours`, `exchange_handles`, `combine_sums`,
`gather_rows`) do not touch the GPU, so CPU tests run them under the gloo backend.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .lattice import SlabLattice, av_from_sums


def _dist():
    import torch.distributed as dist

    return dist


def _device_for_collectives():
    import torch

    dist = _dist()
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def ring_neighbours(rank: int, nranks: int):
    """(south, north) ranks of the periodic ring: up=(r-1+P)%P, down=(r+1)%P (MPI/d2q9-bgk.c:253-254)."""
    return (rank - 1 + nranks) % nranks, (rank + 1) % nranks


def exchange_handles(handle: bytes, group=None):
    """All-gather every rank's halo handle; returns (south_handle, north_handle) of this rank."""
    import torch

    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _device_for_collectives()
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    assert mine.numel() == capi.HALO_HANDLE_BYTES
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    south, north = ring_neighbours(rank, world)
    return bytes(allh[south].cpu().numpy().tobytes()), bytes(allh[north].cpu().numpy().tobytes())


def combine_sums(sums: np.ndarray, nonfinite: np.ndarray, fluid_cells: int, group=None) -> np.ndarray:
    """Add the per-rank integer |u| sums over the group and turn them into av_vels.  Integer addition:
    the result is independent of the number of ranks and of their order (the reference's float
    MPI_Reduce is not, MPI/d2q9-bgk.c:298-309)."""
    import torch

    dist = _dist()
    dev = _device_for_collectives()
    n = sums.shape[0]
    packed = np.concatenate([sums.reshape(-1), nonfinite.reshape(-1), np.array([fluid_cells], dtype=np.int64)])
    t = torch.from_numpy(packed.astype(np.int64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tot = t.cpu().numpy()
    lo_hi = tot[: 2 * n].reshape(n, 2)
    bad = tot[2 * n: 3 * n]
    fluid = int(tot[3 * n])
    return np.array([av_from_sums(lo_hi[i, 0], lo_hi[i, 1], bad[i], fluid) for i in range(n)], dtype=np.float32)


def add_over_ranks(values: np.ndarray, group=None) -> np.ndarray:
    """int64 all-reduce(SUM) of a small host array (the integer |u| sums, fluid-cell counts)."""
    import torch

    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int64)).to(_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def gather_rows(rows: np.ndarray, starts, dst: int = 0, group=None):
    """Gather row slabs to rank `dst` in rank order (MPI/d2q9-bgk.c:265-295).  rows: float32[my_rows, ...].
    Returns the full array on `dst`, None elsewhere."""
    import torch

    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _device_for_collectives()
    tail = rows.shape[1:]
    maxrows = max(starts[r + 1] - starts[r] for r in range(world))
    pad = np.zeros((maxrows,) + tail, dtype=rows.dtype)
    pad[: rows.shape[0]] = rows
    mine = torch.from_numpy(pad).to(dev)
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine, group=group)
    if rank != dst:
        return None
    out = np.empty((starts[world],) + tail, dtype=rows.dtype)
    for r in range(world):
        out[starts[r]: starts[r + 1]] = allr[r].cpu().numpy()[: starts[r + 1] - starts[r]]
    return out


class ShardedLattice:
    """This rank's slab of an nx x ny grid decomposed over the process group, on CUDA device `device`."""

    def __init__(self, param, obstacle_rows_fn, device: int, group=None, **options):
        dist = _dist()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.param = param
        self.starts = capi.partition(param.ny, self.world) if self.world > 1 else [0, param.ny]
        r0, r1 = self.starts[self.rank], self.starts[self.rank + 1]
        rows = obstacle_rows_fn(r0, r1)  # int32[r1-r0, nx]: each rank builds (or reads) only its own rows
        self.slab = SlabLattice(param, rows, r0, r1, self.rank, self.world, device, **options)
        south, north = exchange_handles(self.slab.export_handle(), group)
        self.slab.connect(south, north)
        dist.barrier(group=group)  # every ring is mapped before anyone's first step stores into it

    def run(self, iters: int) -> None:
        self.slab.run(iters)

    def sync(self) -> None:
        self.slab.sync()

    def av_vels(self) -> np.ndarray:
        sums, bad = self.slab.tot_u_sums()
        return combine_sums(sums, bad, self.slab.fluid_cells, self.group)

    def tot_u_totals(self) -> np.ndarray:
        """uint64[iters]: the exact integer |u| total (units of 2^-40, modulo 2^64) of every step of the last run,
        added over the ranks -- equal, bit for bit, to what ONE GPU computes for the same grid."""
        sums, _ = self.slab.tot_u_sums()
        tot = add_over_ranks(sums, self.group).astype(np.uint64)
        return tot[:, 0] + (tot[:, 1] << np.uint64(24))

    def final_state(self, dst: int = 0):
        parts = self.slab.final_state()
        full = [gather_rows(p, self.starts, dst, self.group) for p in parts]
        return tuple(full) if self.rank == dst else None

    def cells(self, dst: int = 0):
        return gather_rows(self.slab.cells(), self.starts, dst, self.group)

    def close(self) -> None:
        _dist().barrier(group=self.group)  # nobody unmaps a ring a neighbour may still store into
        self.slab.close()
