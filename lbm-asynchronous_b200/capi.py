"""ctypes binding of include/lbm_b200.h (one Python function per exported symbol, same names)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")  # override: tuning builds (tools/variants.sh)

NSPEEDS = 9
HALO_HANDLE_BYTES = 128
ARITH_STRICT, ARITH_FAST = 0, 1
HALO_SYNC, HALO_ASYNC = 0, 1
OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ETIMEOUT = 0, 1, 2, 3, 4, 5

# every symbol include/lbm_b200.h declares (tests check the .so exports each of them)
SYMBOLS = (
    "lbm_default_options", "lbm_last_error", "lbm_device_count", "lbm_partition", "lbm_create", "lbm_create_on",
    "lbm_create_packed", "lbm_packed_words_per_row", "lbm_create_slab", "lbm_create_slab_packed", "lbm_halo_export", "lbm_halo_connect", "lbm_set_stream", "lbm_run", "lbm_sync", "lbm_av_vels",
    "lbm_tot_u_sums", "lbm_av_from_sums", "lbm_fluid_cells", "lbm_steps_done", "lbm_av_velocity", "lbm_total_density",
    "lbm_final_state", "lbm_download_cells", "lbm_upload_cells", "lbm_last_run_ms", "lbm_kernel_launches",
    "lbm_num_slabs", "lbm_slab_info", "lbm_destroy", "lbm_selftest", "lbm_selftest_collide",
)


class Param(C.Structure):
    """lbm_param_t == the reference's t_param (SerialCode/d2q9-bgk.c:66-75)."""

    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("maxIters", C.c_int), ("reynolds_dim", C.c_int),
        ("density", C.c_float), ("accel", C.c_float), ("omega", C.c_float),
    ]


class Options(C.Structure):
    _fields_ = [
        ("arith", C.c_int), ("halo_mode", C.c_int), ("halo_lag", C.c_int), ("use_graph", C.c_int),
        ("kernel", C.c_int), ("block", C.c_int),
    ]


class LbmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[LBM error {code}] {message}")
        self.code = code
        self.message = message


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile liblbm_b200.so / d2q9-bgk for sm_100a with the repository Makefile (nvcc
    cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", _ROOT, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", _ROOT, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building liblbm_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return _LIB


_lib = None


def library() -> C.CDLL:
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise LbmError(ENODEVICE, f"{_LIB} is missing: run `make` (or __graft_entry__.build()) first; "
                                  "this package has no CPU or PyTorch fallback")
    L = C.CDLL(_LIB)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    llp = C.POINTER(C.c_longlong)
    pp, op = C.POINTER(Param), C.POINTER(Options)
    sig = {
        "lbm_default_options": (None, [op]),
        "lbm_last_error": (C.c_char_p, []),
        "lbm_device_count": (C.c_int, []),
        "lbm_partition": (C.c_int, [C.c_int, C.c_int, ip]),
        "lbm_create": (C.c_int, [pp, ip, C.c_int, op, C.POINTER(vp)]),
        "lbm_create_on": (C.c_int, [pp, ip, C.c_int, ip, op, C.POINTER(vp)]),
        "lbm_create_slab": (C.c_int, [pp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, op, C.POINTER(vp)]),
        "lbm_create_packed": (C.c_int, [pp, C.POINTER(C.c_uint), C.c_int, op, C.POINTER(vp)]),
        "lbm_packed_words_per_row": (C.c_size_t, [C.c_int]),
        "lbm_create_slab_packed": (C.c_int, [pp, C.POINTER(C.c_uint), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, op, C.POINTER(vp)]),
        "lbm_halo_export": (C.c_int, [vp, vp]),
        "lbm_halo_connect": (C.c_int, [vp, vp, vp]),
        "lbm_set_stream": (C.c_int, [vp, vp]),
        "lbm_run": (C.c_int, [vp, C.c_int]),
        "lbm_sync": (C.c_int, [vp]),
        "lbm_av_vels": (C.c_int, [vp, fp, C.c_int]),
        "lbm_tot_u_sums": (C.c_int, [vp, llp, llp, C.c_int]),
        "lbm_av_from_sums": (C.c_float, [C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong]),
        "lbm_fluid_cells": (C.c_longlong, [vp]),
        "lbm_steps_done": (C.c_longlong, [vp]),
        "lbm_av_velocity": (C.c_int, [vp, fp]),
        "lbm_total_density": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "lbm_final_state": (C.c_int, [vp, fp, fp, fp, fp]),
        "lbm_download_cells": (C.c_int, [vp, fp]),
        "lbm_upload_cells": (C.c_int, [vp, fp]),
        "lbm_last_run_ms": (C.c_int, [vp, fp]),
        "lbm_kernel_launches": (C.c_longlong, [vp]),
        "lbm_num_slabs": (C.c_int, [vp]),
        "lbm_slab_info": (C.c_int, [vp, C.c_int, ip, ip, ip]),
        "lbm_destroy": (None, [vp]),
        "lbm_selftest": (C.c_int, [C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
        "lbm_selftest_collide": (C.c_int, [C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise LbmError(rc, library().lbm_last_error().decode("utf-8", "replace"))


def selftest(pairs: int = 1 << 30, seed: int = 1, device: int = 0):
    """lbm_selftest: (differing quotients, differing roots) over ~`pairs` random operand sets; both must be 0."""
    out = (C.c_ulonglong * 2)()
    check(library().lbm_selftest(device, pairs, seed, out))
    return int(out[0]), int(out[1])


def selftest_collide(arith: str = "strict", sets: int = 1 << 24, seed: int = 1, device: int = 0):
    """lbm_selftest_collide: (differing population words, differing |u| words) of the packed four-cell collision
    against the scalar per-cell code over ~`sets` random groups of four cells."""
    out = (C.c_ulonglong * 2)()
    check(library().lbm_selftest_collide(device, {"strict": 0, "fast": 1}[arith], sets, seed, out))
    return int(out[0]), int(out[1])


def partition(ny: int, nslabs: int):
    """lbm_partition: balanced row slabs, slab r owns rows [starts[r], starts[r+1])."""
    starts = (C.c_int * (nslabs + 1))()
    check(library().lbm_partition(ny, nslabs, starts))
    return list(starts)
