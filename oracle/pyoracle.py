"""ctypes front-end of oracle/lbm_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module (as the checker / the timed CPU baseline).  The product (liblbm_b200.so and the
d2q9-bgk host program) never does.

Also holds small numpy readers for the reference's text formats (params file, obstacle file,
final_state.dat, av_vels.dat; SerialCode/d2q9-bgk.c:480-506, :588-601, :722, :737) and a
restatement of check/check.py's comparison metric (check/check.py:83-99,136-139) so that GPU-box
tests, which cannot see /root/reference, apply the same acceptance rule.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

# libgomp reads its environment when it is first loaded.  Without a binding policy the threads of the
# oracle's OpenMP loops were observed to share ONE core in the build container (8 threads, 1x speed);
# the reference's own env.sh binds threads as well (OpenMP/env.sh:2-4).
os.environ.setdefault("OMP_PROC_BIND", "spread")
os.environ.setdefault("OMP_PLACES", "cores")

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblbm_oracle.so")
NSPEEDS = 9


class OracleParam(C.Structure):
    _fields_ = [
        ("nx", C.c_int),
        ("ny", C.c_int),
        ("max_iters", C.c_int),
        ("reynolds_dim", C.c_int),
        ("density", C.c_float),
        ("accel", C.c_float),
        ("omega", C.c_float),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc only; also builds oracle/_ref when the
    reference tree is mounted)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("lbm_oracle.c", "lbm_oracle.h")
    )
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        pp = C.POINTER(OracleParam)
        L.oracle_init_cells.argtypes = [pp, fp]
        L.oracle_accelerate_flow.argtypes = [pp, fp, ip]
        L.oracle_propagate.argtypes = [pp, fp, fp]
        L.oracle_rebound.argtypes = [pp, fp, fp, ip]
        L.oracle_collision.argtypes = [pp, fp, fp, ip]
        L.oracle_timestep.argtypes = [pp, fp, fp, ip]
        L.oracle_av_velocity.argtypes = [pp, fp, ip]
        L.oracle_av_velocity.restype = C.c_float
        L.oracle_tot_u_f64.argtypes = [pp, fp, ip, ip]
        L.oracle_tot_u_f64.restype = C.c_double
        L.oracle_total_density.argtypes = [pp, fp]
        L.oracle_total_density.restype = C.c_float
        L.oracle_calc_reynolds.argtypes = [pp, fp, ip]
        L.oracle_calc_reynolds.restype = C.c_float
        L.oracle_run.argtypes = [pp, fp, fp, ip, C.c_int, fp]
        L.oracle_final_state.argtypes = [pp, fp, ip, fp, fp, fp, fp]
        L.oracle_fused_step.argtypes = [pp, fp, fp, ip]
        L.oracle_fused_step.restype = C.c_float
        L.oracle_reference_partition.argtypes = [C.c_int, C.c_int, ip]
        L.oracle_reference_partition.restype = C.c_int
        L.oracle_run_decomposed.argtypes = [pp, ip, C.c_int, ip, C.c_int, C.c_int, fp, fp]
        L.oracle_run_decomposed.restype = C.c_int
        for name in (
            "oracle_init_cells",
            "oracle_accelerate_flow",
            "oracle_propagate",
            "oracle_rebound",
            "oracle_collision",
            "oracle_timestep",
            "oracle_run",
            "oracle_final_state",
        ):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _f(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Params:
    """The seven values of the reference's params file (SerialCode/d2q9-bgk.c:480-506)."""

    def __init__(self, nx, ny, max_iters, reynolds_dim, density, accel, omega):
        self.nx, self.ny, self.max_iters, self.reynolds_dim = int(nx), int(ny), int(max_iters), int(reynolds_dim)
        self.density, self.accel, self.omega = np.float32(density), np.float32(accel), np.float32(omega)

    def c(self) -> OracleParam:
        return OracleParam(self.nx, self.ny, self.max_iters, self.reynolds_dim, float(self.density),
                           float(self.accel), float(self.omega))

    def replace(self, **kw) -> "Params":
        d = dict(nx=self.nx, ny=self.ny, max_iters=self.max_iters, reynolds_dim=self.reynolds_dim,
                 density=self.density, accel=self.accel, omega=self.omega)
        d.update(kw)
        return Params(**d)


def read_params(path: str) -> Params:
    with open(path) as fh:
        tok = fh.read().split()
    return Params(int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3]), float(tok[4]), float(tok[5]), float(tok[6]))


def read_obstacles(path: str, nx: int, ny: int) -> np.ndarray:
    """`x y 1` lines -> int32[ny, nx] (SerialCode/d2q9-bgk.c:588-601)."""
    obst = np.zeros((ny, nx), dtype=np.int32)
    data = np.loadtxt(path, dtype=np.int64, ndmin=2)
    if data.size:
        assert data.shape[1] == 3 and np.all(data[:, 2] == 1)
        obst[data[:, 1], data[:, 0]] = 1
    return obst


def read_av_vels(path: str) -> np.ndarray:
    return np.loadtxt(path, usecols=[1])


def read_final_state(path: str) -> np.ndarray:
    """columns: ii jj u_x u_y u pressure obstacle (SerialCode/d2q9-bgk.c:722)."""
    return np.loadtxt(path)


def check_metric(ref: np.ndarray, sim: np.ndarray) -> float:
    """Worst percentage difference exactly as check/check.py:83-99 computes it:
    diff = ref - sim; pcnt = 100*diff/(ref - diff) = 100*(ref-sim)/sim; returns the signed
    value of largest magnitude (nan/inf propagate and must be treated as failure, :136-139)."""
    ref = np.asarray(ref, dtype=np.float64).ravel()
    sim = np.asarray(sim, dtype=np.float64).ravel()
    diff = ref - sim
    with np.errstate(divide="ignore", invalid="ignore"):
        pcnt = 100.0 * (diff / (ref - diff))
    k = int(np.argmax(np.abs(pcnt)))
    return float(pcnt[k])


def check_passes(ref_av, sim_av, ref_pressure, sim_pressure, tolerance: float = 1.0):
    """check.py's verdict (default tolerance 1 %, check/check.py:19-24)."""
    a = check_metric(ref_av, sim_av)
    f = check_metric(ref_pressure, sim_pressure)
    ok = np.isfinite(a) and np.isfinite(f) and abs(a) <= tolerance and abs(f) <= tolerance
    return bool(ok), a, f


# ---------------------------------------------------------------------------------------------
# numpy-level wrappers
# ---------------------------------------------------------------------------------------------

def init_cells(p: Params) -> np.ndarray:
    cells = np.empty((p.ny, p.nx, NSPEEDS), dtype=np.float32)
    cp = p.c()
    lib().oracle_init_cells(C.byref(cp), _f(cells))
    return cells


def run(p: Params, obstacles: np.ndarray, iters: int, cells: np.ndarray | None = None):
    """iters x {timestep; av_velocity} from `cells` (default: the uniform initial state).
    Returns (cells_after, av_vels)."""
    cells = init_cells(p) if cells is None else np.ascontiguousarray(cells, dtype=np.float32).copy()
    tmp = np.zeros_like(cells)
    av = np.zeros(iters, dtype=np.float32)
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    cp = p.c()
    lib().oracle_run(C.byref(cp), _f(cells), _f(tmp), _i(ob), iters, _f(av))
    return cells, av


def run_fused(p: Params, obstacles: np.ndarray, iters: int, cells: np.ndarray | None = None):
    """Same through the fused OpenMP-style pass (ping-pong lattices)."""
    a = init_cells(p) if cells is None else np.ascontiguousarray(cells, dtype=np.float32).copy()
    b = a.copy()
    av = np.zeros(iters, dtype=np.float32)
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    cp = p.c()
    for tt in range(iters):
        av[tt] = lib().oracle_fused_step(C.byref(cp), _f(a), _f(b), _i(ob))
        a, b = b, a
    return a, av


def final_state(p: Params, cells: np.ndarray, obstacles: np.ndarray):
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    outs = [np.empty((p.ny, p.nx), dtype=np.float32) for _ in range(4)]
    cp = p.c()
    lib().oracle_final_state(C.byref(cp), _f(np.ascontiguousarray(cells)), _i(ob), *[_f(o) for o in outs])
    return tuple(outs)  # u_x, u_y, u, pressure


def av_velocity(p: Params, cells: np.ndarray, obstacles: np.ndarray) -> np.float32:
    cp = p.c()
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    return np.float32(lib().oracle_av_velocity(C.byref(cp), _f(np.ascontiguousarray(cells)), _i(ob)))


def tot_u_f64(p: Params, cells: np.ndarray, obstacles: np.ndarray):
    cp = p.c()
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    n = C.c_int(0)
    tot = lib().oracle_tot_u_f64(C.byref(cp), _f(np.ascontiguousarray(cells)), _i(ob), C.byref(n))
    return float(tot), int(n.value)


def total_density(p: Params, cells: np.ndarray) -> np.float32:
    cp = p.c()
    return np.float32(lib().oracle_total_density(C.byref(cp), _f(np.ascontiguousarray(cells))))


def calc_reynolds(p: Params, cells: np.ndarray, obstacles: np.ndarray) -> np.float32:
    cp = p.c()
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    return np.float32(lib().oracle_calc_reynolds(C.byref(cp), _f(np.ascontiguousarray(cells)), _i(ob)))


def reference_partition(ny: int, nranks: int) -> np.ndarray:
    starts = np.zeros(nranks + 1, dtype=np.int32)
    rc = lib().oracle_reference_partition(ny, nranks, _i(starts))
    if rc != 0:
        raise ValueError("reference partition leaves a rank without rows")
    return starts


def run_decomposed(p: Params, obstacles: np.ndarray, starts, halo_lag: int, iters: int):
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    nranks = len(starts) - 1
    cells = np.empty((p.ny, p.nx, NSPEEDS), dtype=np.float32)
    av = np.zeros(iters, dtype=np.float32)
    ob = np.ascontiguousarray(obstacles, dtype=np.int32)
    cp = p.c()
    rc = lib().oracle_run_decomposed(C.byref(cp), _i(ob), nranks, _i(starts), halo_lag, iters, _f(cells), _f(av))
    if rc != 0:
        raise ValueError(f"oracle_run_decomposed rejected its arguments (rc={rc})")
    return cells, av


# ---------------------------------------------------------------------------------------------
# the reference's own programs (oracle/_ref, built by oracle/Makefile from /root/reference)
# ---------------------------------------------------------------------------------------------
REF_DIR = os.path.join(_HERE, "_ref")


def reference_binary(name: str) -> str | None:
    """Path of a prebuilt reference program (serial, openmp, MPI, MPI_Waitall,
    MPI_Testall_OptimizedVersion, each MPI one also with the suffix -strict), or None."""
    path = os.path.join(REF_DIR, "d2q9-bgk-" + name)
    return path if os.path.exists(path) else None


def run_reference(name: str, params_file: str, obstacles_file: str, workdir: str, nranks: int = 1, threads: int | None = None,
                  eager_bytes: int | None = None, discard_final_state: bool = False, timeout: float = 3600.0):
    """Run a reference program in `workdir` (it writes av_vels.dat / final_state.dat there).  MPI programs
    run over oracle/minimpi with `nranks` ranks.  Returns the program's own 'Elapsed Compute time' (s)."""
    import re
    import resource

    exe = reference_binary(name)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/d2q9-bgk-{name} is not built (run `make -C oracle ref` where /root/reference is mounted)")
    os.makedirs(workdir, exist_ok=True)
    if discard_final_state:  # 87 bytes of text per cell: send it to /dev/null for timing runs
        fs = os.path.join(workdir, "final_state.dat")
        if os.path.lexists(fs):
            os.remove(fs)
        os.symlink("/dev/null", fs)
    env = dict(os.environ, MINIMPI_NP=str(nranks))
    if threads:
        env.update(OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="true", OMP_PLACES="cores")  # OpenMP/env.sh:2-4
    if eager_bytes:
        env["MINIMPI_EAGER_BYTES"] = str(eager_bytes)

    def unlimited_stack():  # the MPI variants keep row-sized VLAs on the stack (MPI/d2q9-bgk.c:298,816)
        try:
            resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
        except (ValueError, OSError):
            pass

    r = subprocess.run([exe, params_file, obstacles_file], cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout,
                       preexec_fn=unlimited_stack)
    if r.returncode != 0:
        raise RuntimeError(f"{name} failed ({r.returncode}): {r.stdout[-500:]} {r.stderr[-500:]}")
    m = re.search(r"Elapsed Compute time:\s+([0-9.]+)", r.stdout)
    return float(m.group(1))
