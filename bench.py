#!/usr/bin/env python3
"""bench.py -- MLUPS of the D2Q9-BGK timestep on N B200s (and the reference's CPU arm beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # N=1: in-process
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]

One "step" = one lattice-Boltzmann timestep over the whole grid (accelerate + propagate + rebound +
collision + the step's av_velocity reduction, one kernel launch per slab).

Workloads (BASELINE.json configs 4 and 5):
  N = 1 : synthetic channel 8192 x 8192, Bernoulli(0.005) obstacles, SplitMix64 seed 42
  N > 1 : weak scaling, 32768 columns x 4096 rows PER GPU (8 GPUs = 32768 x 32768), same generator;
          row slabs, one process per GPU; the per-step halo exchange is done by the step kernel
          itself (peer stores into the neighbour GPU's halo ring over NVLink) -- no NCCL on the data
          path; torch.distributed only carries handles, barriers and the timing reduction.

Prints ONE JSON line (rank 0).  `value` = cells * K / device time (CUDA events, max over ranks) with
everything resident in HBM; `e2e` = the same metric through the public API from HOST buffers
(obstacle map uploaded, lattice initialised, K steps, av_vels and the final-state moments downloaded
to pinned host memory inside the timed region; the pass is made twice, the faster one is reported and both
wall times are listed in `e2e.passes_s`).  Timing hygiene: W >= 3 warm-up steps; the two
lattices (4.8 GB at 8192^2) are far larger than L2 (126 MB), so every step streams from HBM.

N > 1 also runs, OUTSIDE the timed region, a parity leg (`parity_check` in the JSON line): the shipped
1024 x 1024 case and a 32768 x (N*64) slab case through the same one-process-per-GPU path, compared
bit for bit (sha256 of every rank's cells, exact integer |u| totals) with a 1-GPU run of the same grid on
rank 0's device -- the cross-GPU halo protocol checked where a driver with N GPUs can see it; a
mismatch in sync mode exits non-zero.  And a second timed run in the stale-halo (async) mode with its
drift against the synchronous run in check.py's metric (`async`), BASELINE config 5's second half.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

BYTES_PER_LUP = 72.0  # 9 fp32 read + 9 fp32 written per lattice update


def measured_peak():
    """HBM GB/s denominator: MEASURED_PEAKS.json (driver-written), else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def known_traffic(nx: int, rows_per_gpu: int, arith: str):
    """DRAM bytes per launch of the dominant kernel from a committed `ncu --set full` capture of THIS workload
    shape (profiles/step_kernel_traffic.json: {"<nx>x<rows>:<arith>": {"dram_bytes_per_launch": ..}}), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as fh:
            return json.load(fh).get(f"{nx}x{rows_per_gpu}:{arith}")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        self._nv = self._h = None
        self._names = {}

    def prepare(self) -> bool:
        """NVML initialised and the handle fetched on the caller's thread, BEFORE the timed region (nvmlInit alone
        can take longer than a short timed region)."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            self._names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            self.ok = True
        except Exception:
            self.ok = False
        return self.ok

    def sample(self):
        nv = self._nv
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for bit, name in self._names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        try:
            while not self._stop_evt.is_set():
                self.sample()
                time.sleep(self.period)
        except Exception:
            self.ok = False

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if self.ok:
            try:
                self.sample()  # one more under load: the region may be shorter than the sampling period
            except Exception:
                pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref: the reference's own OpenMP program, built from
# /root/reference by oracle/Makefile) on a bounded sample of the workload
# -------------------------------------------------------------------------------------------------
def run_reference_cpu(nx: int, rows: int, iters: int, threads: int | None = None, program: str = "openmp"):
    """Runs a prebuilt reference program (oracle/_ref: the reference's own sources compiled by oracle/Makefile;
    the MPI programs over oracle/minimpi with one rank per core) on an nx x rows channel of the benchmark's
    generator for `iters` steps; returns (MLUPS from the program's own 'Elapsed Compute time', seconds, kind).
    Falls back to the oracle's fused port when the prebuilt binary is absent."""
    threads = threads or os.cpu_count() or 1
    orc = entry.load_oracle()
    gen = os.path.join(ROOT, "lbm-asynchronous_b200", "gen_channel")
    if orc.reference_binary(program) and os.path.exists(gen):
        td = tempfile.mkdtemp(prefix="lbm_ref_")
        try:
            subprocess.run([gen, str(nx), str(rows), str(iters), "in.params", "in.obstacles"], cwd=td, check=True)
            mpi = program != "openmp" and program != "serial"
            nranks = min(threads, max(1, (rows - 3))) if mpi else 1
            secs = orc.run_reference(program, os.path.join(td, "in.params"), os.path.join(td, "in.obstacles"), td, nranks=nranks,
                                     threads=None if mpi else threads, discard_final_state=True)
            return nx * rows * iters / secs / 1e6, secs, "reference"
        finally:
            shutil.rmtree(td, ignore_errors=True)
    pkg = entry.load_package()
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    p = orc.Params(nx, rows, iters, 10, 0.1, 0.005, 1.85)
    obst = pkg.channel_obstacles(nx, rows)
    a = orc.init_cells(p)
    t0 = time.perf_counter()
    orc.run_fused(p, obst, iters, cells=a)
    secs = time.perf_counter() - t0
    return nx * rows * iters / secs / 1e6, secs, "port"


REF_MARCH = "x86-64-v3 (oracle/Makefile; the reference's Makefiles say -march=native, replaced so that the binary built in the build container runs on the GPU box's host CPU)"


def shipped_grids_cpu(cores: int):
    """The reference's own OpenMP program (oracle/_ref) on the four shipped grids at their FULL iteration counts,
    all host cores, final_state.dat sent to /dev/null: seconds and MLUPS from its 'Elapsed Compute time' line."""
    orc = entry.load_oracle()
    out = {}
    if not orc.reference_binary("openmp"):
        return out
    gin = os.path.join(ROOT, "tests", "golden", "inputs")
    for g in ("128x128", "128x256", "256x256", "1024x1024"):
        td = tempfile.mkdtemp(prefix="lbm_ship_")
        try:
            pf, of = os.path.join(gin, f"input_{g}.params"), os.path.join(gin, f"obstacles_{g}.dat")
            with open(pf) as fh:
                tok = fh.read().split()
            nx, ny, iters = int(tok[0]), int(tok[1]), int(tok[2])
            secs = orc.run_reference("openmp", pf, of, td, threads=cores, discard_final_state=True, timeout=900)
            out[g] = {"seconds": round(secs, 3), "mlups": round(nx * ny * iters / secs / 1e6, 1), "iters": iters}
        except Exception as ex:
            out[g] = {"failed": str(ex)[:120]}
        finally:
            shutil.rmtree(td, ignore_errors=True)
    return out


def shipped_grids_gpu(pkg, arith: str):
    """The same four shipped cases at full iteration counts on ONE GPU through the library (device time of lbm_run)."""
    gin = os.path.join(ROOT, "tests", "golden", "inputs")
    out = {}
    for g in ("128x128", "128x256", "256x256", "1024x1024"):
        p = pkg.read_params(os.path.join(gin, f"input_{g}.params"))
        ob = pkg.read_obstacles(os.path.join(gin, f"obstacles_{g}.dat"), p.nx, p.ny)
        with pkg.Lattice(p, ob, ngpus=1, arith=arith) as lat:
            lat.run(p.maxIters)
            ms = lat.last_run_ms()
        out[g] = {"seconds": round(ms * 1e-3, 4), "mlups": round(p.nx * p.ny * p.maxIters / (ms * 1e-3) / 1e6, 1),
                  "us_per_step": round(ms * 1e3 / p.maxIters, 3)}
    return out


def reference_arm(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nx = args.nx or (8192 if args.gpus == 1 else 32768)
    ny = (args.ny or (8192 if args.gpus == 1 else 4096)) * (1 if args.gpus == 1 else args.gpus)
    cores = os.cpu_count() or 1
    # bounded sample: rows chosen so that (warm-up + K) steps are roughly 4e9 lattice updates
    total_steps = args.steps + args.warmup
    rows = int(4.0e9 / (nx * max(total_steps, 1)))
    rows = max(64, min(2048, rows // 64 * 64))
    rows = max(64, min(rows, (1 << 24) // nx))  # the reference formats 87 B of text per cell at the end: <= 16 M cells
    run_reference_cpu(nx, rows, max(args.warmup, 1), cores)  # warm-up run (page cache, CPU clocks)
    # every maintained variant of the reference on the same sample; the arm's value is the fastest one
    results = {}
    for prog in ("openmp", "MPI_Waitall", "MPI_Testall_OptimizedVersion", "MPI"):
        try:
            results[prog] = run_reference_cpu(nx, rows, args.steps, cores, program=prog)
        except Exception as ex:
            print(f"bench.py: reference program {prog} failed: {ex}", file=sys.stderr)
    if not results:  # no prebuilt reference program ran on this box: time the oracle's restatement instead
        results["oracle port (fused OpenMP pass)"] = run_reference_cpu(nx, rows, args.steps, cores, program="__none__")
    best = max(results, key=lambda k: results[k][0])
    mlups, secs, kind = results[best]
    sample = (f"{nx}x{rows} rows of the channel workload, {args.steps} steps, {cores} host threads/ranks; fastest reference "
              f"program = {best}; all: " + ", ".join(f"{k} {v[0]:.0f}" for k, v in results.items()) + " MLUPS")
    cfg = workload_config(args.gpus, nx, ny)
    cfg["workload"] = (f"SAMPLE {nx}x{rows} rows ({args.steps} steps) of: " + cfg["workload"] +
                       " -- MLUPS of a row sample (the CPU programs' MLUPS does not depend on the row count at this size)")
    cfg["sample_nx"], cfg["sample_rows"] = nx, rows
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample, "march": REF_MARCH},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(index: int):
    """Run this rank's host threads (and so its pinned allocations, first touch) on the NUMA node the GPU hangs
    off: pinned buffers on the far socket cost the final-state download a hop over the inter-socket link.
    Returns a short description, or None when the topology cannot be read."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} cpus"
    except Exception:
        return None


def workload_config(n_gpus: int, nx: int, ny: int | None, note: str | None = None):
    if n_gpus == 1:
        name = f"synthetic channel {nx}x{ny or 8192}, Bernoulli(0.005) obstacles, SplitMix64 seed 42 (BASELINE config 4)"
    else:
        name = (f"synthetic channel weak scaling {nx} x {ny} ({nx}x{(ny or 0) // n_gpus} per GPU, row slabs), "
                "Bernoulli(0.005) obstacles, seed 42 (BASELINE config 5)")
    c = {"workload": name, "nx": nx, "ny": ny, "physics": "reynolds_dim=10 density=0.1 accel=0.005 omega=1.85",
         "cache": "inputs larger than L2 (two lattices of 36 B/cell each; 126 MB L2)"}
    if note:
        c["note"] = note
    return c


# -------------------------------------------------------------------------------------------------
# N > 1: cross-GPU parity leg (outside the timed region)
# -------------------------------------------------------------------------------------------------
def parity_leg(pkg, dist, rank, world, local_rank, opts):
    """Same one-process-per-GPU path as the timed run, on two grids small enough for a 1-GPU re-run on rank 0:
    sync mode must equal the 1-GPU lattice bit for bit (sha256 per rank slab) and in the exact integer |u| totals;
    async mode reports its drift in check.py's metric.  Reference semantics: MPI_Waitall/d2q9-bgk.c:225-253
    (sync), MPI_Testall_OptimizedVersion/d2q9-bgk.c:263-290 (async)."""
    import torch

    from lbm_asynchronous_b200.lattice import make_param
    from lbm_asynchronous_b200.sharded import ShardedLattice

    gin = os.path.join(ROOT, "tests", "golden", "inputs")
    p1024 = pkg.read_params(os.path.join(gin, "input_1024x1024.params"))
    ob1024 = pkg.read_obstacles(os.path.join(gin, "obstacles_1024x1024.dat"), 1024, 1024)
    wide_rows = 64 * world
    cases = [
        ("shipped 1024x1024", p1024, lambda a, b: ob1024[a:b], [150, 151]),
        (f"synthetic channel 32768x{wide_rows}", make_param(32768, wide_rows, 0), lambda a, b: pkg.channel_obstacles(32768, wide_rows, row0=a, row1=b), [50]),
    ]
    report = {"cases": [], "bit_identical": True, "async_drift_pct": {}}
    sync_opts = dict(opts, halo_mode="sync")
    for name, param, obst_fn, runs in cases:
        def run_sharded(o):
            sh = ShardedLattice(param, obst_fn, local_rank, **o)
            for k in runs:
                sh.run(k)
            sh.sync()
            cells = sh.slab.cells()
            digest = hashlib.sha256(cells.tobytes()).digest()
            totals = sh.tot_u_totals()
            pr = sh.slab.pressure()
            av = sh.av_vels()
            starts = list(sh.starts)
            sh.close()
            return digest, totals, pr, av, starts

        digest, totals, pr_sync, av_sync, starts = run_sharded(sync_opts)
        mine = torch.frombuffer(bytearray(digest), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ok = True
        if rank == 0:
            with pkg.Lattice(param, obst_fn(0, param.ny), ngpus=1, **{k: v for k, v in sync_opts.items() if k != "halo_mode"}) as one:
                for k in runs:
                    one.run(k)
                cells1 = one.cells()
                s1, _ = one.tot_u_sums()
            t1 = s1[:, 0].astype(np.uint64) + (s1[:, 1].astype(np.uint64) << np.uint64(24))
            slabs_equal = [hashlib.sha256(np.ascontiguousarray(cells1[starts[r]:starts[r + 1]]).tobytes()).digest() ==
                           bytes(allh[r].cpu().numpy().tobytes()) for r in range(world)]
            totals_equal = bool(np.array_equal(t1, totals))
            ok = all(slabs_equal) and totals_equal
            report["cases"].append({"grid": name, "steps": int(sum(runs)), "runs": runs, "mode": "sync", "slabs_bit_identical": slabs_equal,
                                    "tot_u_integer_totals_equal": totals_equal})
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        if not int(flag.item()):
            report["bit_identical"] = False
        # stale-halo mode on the same case: drift against the synchronous run, check.py's metric
        _, _, pr_async, av_async, _ = run_sharded(dict(opts, halo_mode="async"))
        d_av = pkg.check_metric(av_sync, av_async)
        d_pr = pkg.check_metric(pr_sync, pr_async)
        t = torch.tensor([abs(d_pr) if np.isfinite(d_pr) else float("inf")], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        report["async_drift_pct"][name] = {"av_vels": d_av, "pressure_max_abs": float(t.item()), "steps": int(sum(runs))}
    return report


# -------------------------------------------------------------------------------------------------
# the GPU arm
# -------------------------------------------------------------------------------------------------
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0, help="rows (N=1) or rows per GPU (N>1)")
    ap.add_argument("--arith", default=os.environ.get("LBM_ARITH", "strict"), choices=["strict", "fast"])
    ap.add_argument("--halo-mode", default="sync", choices=["sync", "async"])
    ap.add_argument("--kernel", type=int, default=int(os.environ.get("LBM_KERNEL", "0")))
    ap.add_argument("--block", type=int, default=int(os.environ.get("LBM_BLOCK", "0")))
    ap.add_argument("--developed", action="store_true",
                    help="N=1 only: start from a perturbed state (every cell moving) instead of the uniform state at rest")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the cross-GPU parity leg")
    ap.add_argument("--no-async", action="store_true", help="N>1: skip the second (stale-halo) timed run")
    ap.add_argument("--no-shipped", action="store_true", help="N=1: skip the shipped-grid CPU/GPU timings")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks", file=sys.stderr)
            return 2
        args.gpus = world
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the product has no CPU path", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = entry.load_package()
    from lbm_asynchronous_b200.lattice import make_param
    from lbm_asynchronous_b200.sharded import ShardedLattice

    n = args.gpus
    nx = args.nx or (8192 if n == 1 else 32768)
    ny = (args.ny or (8192 if n == 1 else 4096)) * (1 if n == 1 else n)
    cells = nx * ny
    K, W = args.steps, args.warmup
    param = make_param(nx, ny, K, 10, 0.1, 0.005, 1.85)
    base_opts = dict(arith=args.arith, kernel=args.kernel, block=args.block)
    opts = dict(base_opts, halo_mode=args.halo_mode)
    # a non-default torch stream: the library runs its kernels on it (lbm_set_stream), the CUDA events
    # below are recorded on it (handle 0, the legacy default stream, means "library's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: cross-GPU parity, before anything is timed ----
    parity = None
    if n > 1 and not args.no_parity:
        parity = parity_leg(pkg, dist, rank, world, local_rank, base_opts)
        if not parity["bit_identical"]:
            if rank == 0:
                print("bench.py: the N-GPU synchronous run differs from the 1-GPU run: " + json.dumps(parity), file=sys.stderr)
            dist.destroy_process_group()
            return 4

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
        sampler.prepare()

    def timed_run(run_opts, want_state):
        """W warm-up steps, then K timed steps (CUDA events on the launching stream, max over ranks)."""
        if n == 1:
            obst = pkg.channel_obstacles(nx, ny)
            lat = pkg.Lattice(param, obst, ngpus=1, **run_opts)
            runner = lat
            if args.developed:
                # every cell gets a small random velocity: no cell takes the "fluid at rest" shortcuts of the kernel
                rng = np.random.default_rng(7)
                cells0 = np.empty((ny, nx, 9), dtype=np.float32)
                w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float32) * np.float32(0.1)
                for k in range(9):
                    cells0[:, :, k] = w[k] * (1 + 0.02 * rng.standard_normal((ny, nx), dtype=np.float32))
                lat.upload(cells0)
                del cells0
        else:
            runner = ShardedLattice(param, lambda r0, r1: pkg.channel_obstacles(nx, ny, row0=r0, row1=r1), local_rank, **run_opts)
            lat = runner.slab
        lat.set_stream(stream.cuda_stream)
        runner.run(W)
        runner.sync()
        barrier()
        launches0 = lat.kernel_launches
        sampling = rank == 0 and sampler.ok and not sampler.is_alive() and not sampler.samples
        if sampling:
            sampler.sample()  # at least one sample exists even if the region is shorter than the sampling period
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        runner.run(K)
        e1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampling else None
        barrier()
        ms = e0.elapsed_time(e1)
        if os.environ.get("LBM_DEBUG"):
            lat.last_run_ms()
        launches = lat.kernel_launches - launches0
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        av = runner.av_vels()
        pressure = lat.pressure() if want_state else None
        lat.set_stream(None)
        runner.close()
        return ms, launches, av, pressure, clocks

    # ---- resident run: `value` ----
    want_async = n > 1 and not args.no_async and args.halo_mode == "sync"
    ms, launches, av, pr_sync, clocks = timed_run(opts, want_async)
    if not np.all(np.isfinite(av)) or not (av > 0).all():
        print("bench.py: av_vels of the timed run are not finite/positive: the run is invalid", file=sys.stderr)
        return 3
    mlups = cells * K / (ms * 1e-3) / 1e6

    # ---- N > 1: the same timed run in the stale-halo mode, drift against the synchronous run ----
    async_line = None
    if want_async:
        ms_a, _, av_a, pr_a, _ = timed_run(dict(base_opts, halo_mode="async"), True)
        d_pr = pkg.check_metric(pr_sync, pr_a)
        t = torch.tensor([abs(d_pr) if np.isfinite(d_pr) else float("inf")], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        async_line = {"value": cells * K / (ms_a * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": ms_a / K,
                      "drift_pct": {"av_vels": pkg.check_metric(av, av_a), "pressure": float(t.item())},
                      "what": f"halo_mode=async (boundary rows never wait, MPI_Testall_OptimizedVersion/d2q9-bgk.c:263-290), same grid, "
                              f"{W}+{K} steps from the same initial state; drift = check.py's worst 100*(sync-async)/async"}
        del pr_a
    del pr_sync

    # ---- end to end through the public API from host buffers: `e2e` ----
    e2e = None
    if not args.no_e2e:
        if n == 1:
            r0, r1 = 0, ny
        else:
            st = pkg.partition(ny, n)
            r0, r1 = st[rank], st[rank + 1]
        my_cells = (r1 - r0) * nx
        numa = bind_to_gpu_numa_node(local_rank)
        # the obstacle map as the host program holds it: one bit per cell (lbm_create_packed / lbm_create_slab_packed)
        words = (nx + 31) // 32
        obst_pinned = torch.empty((r1 - r0, words), dtype=torch.int32, pin_memory=True)
        obst_pinned.numpy().view(np.uint32)[:] = pkg.pack_obstacles(pkg.channel_obstacles(nx, ny, row0=r0, row1=r1))
        obst_packed = obst_pinned.numpy().view(np.uint32)
        outs = [torch.empty((r1 - r0, nx), dtype=torch.float32, pin_memory=True) for _ in range(4)]
        # the whole pass (create + upload, K steps, av_vels, final state to pinned host) twice; the faster pass is the
        # line's value, both are listed: lattice creation instantiates CUDA graphs, which takes anything from
        # milliseconds to a third of a second on a busy host (profiles/r01_variance.md)
        e2e_passes = []
        for _pass in range(2):
            barrier()
            t0 = time.perf_counter()
            if n == 1:
                lat2 = pkg.Lattice(param, obst_packed, ngpus=1, **opts)
                run2 = lat2
            else:
                run2 = ShardedLattice(param, lambda a, b: obst_packed, local_rank, **opts)
                lat2 = run2.slab
            lat2.sync()
            t1 = time.perf_counter()
            run2.run(K)
            lat2.sync()
            t2 = time.perf_counter()
            av2 = run2.av_vels()
            t3 = time.perf_counter()
            import ctypes as C

            from lbm_asynchronous_b200.capi import check, library

            check(library().lbm_final_state(lat2._h, *[C.cast(o.data_ptr(), C.POINTER(C.c_float)) for o in outs]))
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            barrier()
            secs = time.perf_counter() - t0
            phases = {"create_upload": t1 - t0, "run": t2 - t1, "av_vels": t3 - t2, "final_state_download": t4 - t3,
                      "final_state_GBps": my_cells * 16 / max(t4 - t3, 1e-9) / 1e9}
            if dist is not None:
                t = torch.tensor([secs], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                secs = float(t.item())
            e2e = {"value": cells * K / secs / 1e6, "unit": "MLUPS",
                   "h2d_bytes_per_step": (r1 - r0) * words * 4 * n / K, "d2h_bytes_per_step": (my_cells * 16 * n + K * 8 * 3) / K,
                   "seconds": secs, "phases_s_rank0": phases, "host_binding_rank0": numa,
                   "what": "lbm_create_packed(host obstacle bit map) + lbm_run(K) + lbm_av_vels + lbm_final_state to pinned host"}
            run2.close()
            e2e_passes.append(e2e)
        e2e = dict(max(e2e_passes, key=lambda d: d["value"]))
        e2e["passes_s"] = [round(d["seconds"], 6) for d in e2e_passes]
        e2e["what"] += "; faster of two passes (passes_s)"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    achieved = mlups * 1e6 * BYTES_PER_LUP / 1e9 / n  # per GPU, GB/s
    traffic = known_traffic(nx, ny // n, args.arith)
    line = {
        "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": n, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(n, nx, ny), arith=args.arith, halo_mode=args.halo_mode if n > 1 else None,
                       kernel=args.kernel, block=args.block, initial_state="perturbed" if args.developed else "uniform at rest"),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_LUP * cells / n * (traffic or {}).get("steps_per_launch", 1),
                     "kernel": (traffic or {}).get("kernel", "lbm::step_tma_kernel"), "per_gpu": True,
                     "frac_of_nominal_8TBps": achieved / 8000.0,
                     "note": "achieved = 72 B (one-pass algorithmic bytes, SURVEY 8d) x lattice updates / time.  `peak` is the driver's "
                             "torch copy_ figure, not the HBM limit: the strict single-step kernel streams more than that copy does, so "
                             "frac can exceed 1 (ncu: 80 % of the DRAM peak sustained); a kernel that advances two timesteps per HBM pass "
                             "(fast flavour) moves ~36 B per update -- `traffic` (ncu dram bytes per launch of this workload shape, null "
                             "when not captured) is the measured figure"},
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": e2e,
    }
    if parity is not None:
        line["parity_check"] = parity
    if async_line is not None:
        line["async"] = async_line
    if n == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rows, iters = 2048, 20
        try:
            v, secs, kind = run_reference_cpu(nx, rows, iters, cores)
            line["cpu_baseline"] = {"value": v, "unit": "MLUPS", "cores": cores, "kind": kind, "seconds": secs, "march": REF_MARCH,
                                    "sample": f"{nx}x{rows} rows of the same channel workload, {iters} steps, OpenMP program of the reference"}
            # the reference's MPI programs on the same sample, one rank per core over oracle/minimpi (reported beside it)
            others = {}
            for prog in ("MPI", "MPI_Waitall", "MPI_Testall_OptimizedVersion"):
                try:
                    mv, ms_, mk = run_reference_cpu(nx, rows, iters, cores, program=prog)
                    if mk == "reference":
                        others[prog] = round(mv, 1)
                except Exception as ex:
                    others[prog] = f"failed: {str(ex)[:80]}"
            line["cpu_baseline"]["mpi_programs_mlups"] = others
            try:  # SerialCode on one core, a smaller sample (it is ~15x slower)
                sv, _, sk = run_reference_cpu(nx, 256, 5, 1, program="serial")
                if sk == "reference":
                    line["cpu_baseline"]["serial_1core_mlups"] = round(sv, 1)
            except Exception as ex:
                line["cpu_baseline"]["serial_1core_mlups"] = f"failed: {str(ex)[:80]}"
            if not args.no_shipped:
                # the shipped grids at full iteration counts, same box, CPU program beside the GPU
                try:
                    line["cpu_baseline"]["shipped"] = {"cpu_openmp": shipped_grids_cpu(cores), "gpu_1xB200": shipped_grids_gpu(pkg, args.arith),
                                                       "what": "full iteration counts of the four shipped cases; CPU = the reference's OpenMP "
                                                               f"program on {cores} host threads (its own 'Elapsed Compute time'), GPU = device time of lbm_run"}
                except Exception as ex:
                    line["cpu_baseline"]["shipped"] = {"failed": str(ex)[:200]}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
