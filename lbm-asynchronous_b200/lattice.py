"""Host-side mirror of the reference's step interface over the C ABI.

The reference exposes free functions over caller-owned arrays (`timestep`, `av_velocity`,
`total_density`, `calc_reynolds`, `write_values`; SerialCode/d2q9-bgk.c:95-128).  `Lattice` keeps
those names and meanings; the arrays live in HBM behind an opaque `lbm_lattice_t`.  Host-visible
layouts are the reference's: cells as float32[ny, nx, 9] (AoS `t_speed`), obstacles as int32[ny, nx].
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import LbmError, Options, Param, check, library


def _fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _options(arith, halo_mode, halo_lag, use_graph, kernel, block) -> Options:
    o = Options()
    library().lbm_default_options(C.byref(o))
    o.arith = {"strict": capi.ARITH_STRICT, "fast": capi.ARITH_FAST}.get(arith, arith)
    o.halo_mode = {"sync": capi.HALO_SYNC, "async": capi.HALO_ASYNC}.get(halo_mode, halo_mode)
    o.halo_lag = int(halo_lag)
    o.use_graph = int(bool(use_graph))
    o.kernel = int(kernel)
    o.block = int(block)
    return o


def pack_obstacles(obstacles: np.ndarray) -> np.ndarray:
    """int map [rows, nx] (non-zero = blocked) -> the packed map lbm_create_packed takes: uint32[rows, ceil(nx/32)],
    cell x is bit x % 32 of word x // 32."""
    ob = np.asarray(obstacles) != 0
    rows, nx = ob.shape
    words = (nx + 31) // 32
    padded = np.zeros((rows, words * 32), dtype=np.uint8)
    padded[:, :nx] = ob
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(rows, words).copy()


def _is_packed(obstacles, nx: int) -> bool:
    a = np.asarray(obstacles)
    return a.dtype == np.uint32 and a.ndim == 2 and a.shape[1] == (nx + 31) // 32 and not (nx == a.shape[1])


def _uptr(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint))


def make_param(nx, ny, maxIters=0, reynolds_dim=10, density=0.1, accel=0.005, omega=1.85) -> Param:
    return Param(int(nx), int(ny), int(maxIters), int(reynolds_dim), float(density), float(accel), float(omega))


def av_from_sums(lo: int, hi: int, nonfinite: int, fluid_cells: int) -> np.float32:
    """lbm_av_from_sums: av_vels value from (rank-added) integer sums."""
    return np.float32(library().lbm_av_from_sums(int(lo), int(hi), int(nonfinite), int(fluid_cells)))


class Lattice:
    """The device-resident lattices of one run on one or several GPUs of this process
    (lbm_create / lbm_create_on)."""

    def __init__(self, param: Param, obstacles: np.ndarray, ngpus: int = 1, devices=None, arith="strict",
                 halo_mode="sync", halo_lag=0, use_graph=True, kernel=0, block=0):
        self.param = param
        self.nx, self.ny = param.nx, param.ny
        opt = _options(arith, halo_mode, halo_lag, use_graph, kernel, block)
        self._h = C.c_void_p()
        self._rows = param.ny
        self._last_iters = 0
        if _is_packed(obstacles, param.nx):  # uint32[ny, ceil(nx/32)]: one bit per cell (pack_obstacles)
            if devices is not None:
                raise ValueError("a packed obstacle map goes through lbm_create_packed (devices 0..ngpus-1)")
            pk = np.ascontiguousarray(obstacles, dtype=np.uint32)
            check(library().lbm_create_packed(C.byref(param), _uptr(pk), int(ngpus), C.byref(opt), C.byref(self._h)))
            return
        ob = np.ascontiguousarray(obstacles, dtype=np.int32).reshape(param.ny, param.nx)
        if devices is not None:
            dev = (C.c_int * len(devices))(*devices)
            check(library().lbm_create_on(C.byref(param), _iptr(ob), len(devices), dev, C.byref(opt), C.byref(self._h)))
        else:
            check(library().lbm_create(C.byref(param), _iptr(ob), int(ngpus), C.byref(opt), C.byref(self._h)))

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            library().lbm_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the hot path ----
    def run(self, iters: int) -> None:
        """`iters` x { timestep(); av_vels[tt] = av_velocity(); }  (SerialCode/d2q9-bgk.c:166-169); asynchronous."""
        check(library().lbm_run(self._h, int(iters)))
        self._last_iters = int(iters)

    def timestep(self) -> None:
        """One timestep() (SerialCode/d2q9-bgk.c:207-214)."""
        self.run(1)

    def sync(self) -> None:
        check(library().lbm_sync(self._h))

    def set_stream(self, cuda_stream: int | None) -> None:
        check(library().lbm_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def av_vels(self, iters: int | None = None) -> np.ndarray:
        """av_vels[tt] of the last run() call."""
        n = self._last_iters if iters is None else int(iters)
        out = np.zeros(max(n, 0), dtype=np.float32)
        if n > 0:
            check(library().lbm_av_vels(self._h, _fptr(out), n))
        return out

    def tot_u_sums(self, iters: int | None = None):
        """(sums int64[iters, 2], nonfinite int64[iters]) of the last run() call (lbm_tot_u_sums)."""
        n = self._last_iters if iters is None else int(iters)
        sums = np.zeros((max(n, 0), 2), dtype=np.int64)
        bad = np.zeros(max(n, 0), dtype=np.int64)
        check(library().lbm_tot_u_sums(self._h, sums.ctypes.data_as(C.POINTER(C.c_longlong)),
                                       bad.ctypes.data_as(C.POINTER(C.c_longlong)), n))
        return sums, bad

    # ---- state queries (the reference's helper functions) ----
    def av_velocity(self) -> np.float32:
        """av_velocity() of the current state (SerialCode/d2q9-bgk.c:409-458)."""
        v = C.c_float()
        check(library().lbm_av_velocity(self._h, C.byref(v)))
        return np.float32(v.value)

    def total_density(self) -> float:
        """total_density() (SerialCode/d2q9-bgk.c:644-660)."""
        v = C.c_double()
        check(library().lbm_total_density(self._h, C.byref(v)))
        return float(v.value)

    def calc_reynolds(self) -> np.float32:
        """calc_reynolds() (SerialCode/d2q9-bgk.c:637-642), fp32 like the reference."""
        one, two, six = np.float32(1.0), np.float32(2.0), np.float32(6.0)
        viscosity = one / six * (two / np.float32(self.param.omega) - one)
        return np.float32(np.float32(self.av_velocity() * np.float32(self.param.reynolds_dim)) / viscosity)

    def final_state(self):
        """(u_x, u_y, u, pressure), float32[rows, nx] each: what write_values() prints (SerialCode:679-724)."""
        outs = [np.empty((self._rows, self.nx), dtype=np.float32) for _ in range(4)]
        check(library().lbm_final_state(self._h, *[_fptr(o) for o in outs]))
        return tuple(outs)

    def pressure(self) -> np.ndarray:
        """Only the pressure column of final_state() (the one check.py compares), float32[rows, nx]."""
        out = np.empty((self._rows, self.nx), dtype=np.float32)
        check(library().lbm_final_state(self._h, None, None, None, _fptr(out)))
        return out

    def cells(self) -> np.ndarray:
        """The lattice as the reference's AoS array, float32[rows, nx, 9]."""
        out = np.empty((self._rows, self.nx, capi.NSPEEDS), dtype=np.float32)
        check(library().lbm_download_cells(self._h, _fptr(out)))
        return out

    def upload(self, cells: np.ndarray) -> None:
        a = np.ascontiguousarray(cells, dtype=np.float32).reshape(self._rows, self.nx, capi.NSPEEDS)
        check(library().lbm_upload_cells(self._h, _fptr(a)))

    # ---- bookkeeping ----
    @property
    def fluid_cells(self) -> int:
        return int(library().lbm_fluid_cells(self._h))

    @property
    def steps_done(self) -> int:
        return int(library().lbm_steps_done(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(library().lbm_kernel_launches(self._h))

    def last_run_ms(self) -> float:
        v = C.c_float()
        check(library().lbm_last_run_ms(self._h, C.byref(v)))
        return float(v.value)

    def slabs(self):
        n = library().lbm_num_slabs(self._h)
        out = []
        for i in range(n):
            r0, r1, d = C.c_int(), C.c_int(), C.c_int()
            check(library().lbm_slab_info(self._h, i, C.byref(r0), C.byref(r1), C.byref(d)))
            out.append((r0.value, r1.value, d.value))
        return out


class SlabLattice(Lattice):
    """One row slab of a grid that is decomposed over several processes, one GPU each
    (lbm_create_slab + lbm_halo_export / lbm_halo_connect)."""

    def __init__(self, param: Param, obstacle_rows: np.ndarray, row0: int, row1: int, rank: int, nranks: int,
                 device: int, arith="strict", halo_mode="sync", halo_lag=0, use_graph=True, kernel=0, block=0):
        self.param = param
        self.nx, self.ny = param.nx, param.ny
        self.row0, self.row1, self.rank, self.nranks = int(row0), int(row1), int(rank), int(nranks)
        self._rows = self.row1 - self.row0
        opt = _options(arith, halo_mode, halo_lag, use_graph, kernel, block)
        self._h = C.c_void_p()
        self._last_iters = 0
        if _is_packed(obstacle_rows, param.nx):
            pk = np.ascontiguousarray(obstacle_rows, dtype=np.uint32)
            check(library().lbm_create_slab_packed(C.byref(param), _uptr(pk), self.row0, self.row1, self.rank, self.nranks,
                                                   int(device), C.byref(opt), C.byref(self._h)))
            return
        ob = np.ascontiguousarray(obstacle_rows, dtype=np.int32).reshape(self._rows, param.nx)
        check(library().lbm_create_slab(C.byref(param), _iptr(ob), self.row0, self.row1, self.rank, self.nranks,
                                        int(device), C.byref(opt), C.byref(self._h)))

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(capi.HALO_HANDLE_BYTES)
        check(library().lbm_halo_export(self._h, buf))
        return bytes(buf.raw)

    def connect(self, south_handle: bytes, north_handle: bytes) -> None:
        s = C.create_string_buffer(bytes(south_handle), capi.HALO_HANDLE_BYTES)
        n = C.create_string_buffer(bytes(north_handle), capi.HALO_HANDLE_BYTES)
        check(library().lbm_halo_connect(self._h, s, n))


__all__ = ["Lattice", "SlabLattice", "LbmError", "make_param", "av_from_sums", "pack_obstacles"]
