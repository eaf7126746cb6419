"""B200-native D2Q9-BGK lattice-Boltzmann timestep (host-side Python mirror of the C ABI).

The product is `liblbm_b200.so` (hand-written sm_100a CUDA behind include/lbm_b200.h) and the
`d2q9-bgk` C host program.  This package is the thin ctypes layer tests, `bench.py` and Python
callers use; every compute call goes through the C ABI.  There is no CPU fallback: creating a
lattice without a CUDA device raises `LbmError` (LBM_ENODEVICE).

The directory name contains a hyphen (it is the repository's package directory,
`lbm-asynchronous_b200/`); `__graft_entry__.load_package()` registers it as the importable module
`lbm_asynchronous_b200`.
"""
from .capi import (  # noqa: F401
    ARITH_FAST,
    ARITH_STRICT,
    HALO_ASYNC,
    HALO_SYNC,
    LbmError,
    Options,
    Param,
    build_library,
    library,
    library_path,
    partition,
    selftest,
    selftest_collide,
)
from .lattice import Lattice, SlabLattice, av_from_sums, pack_obstacles  # noqa: F401
from .inputs import check_metric, read_obstacles, read_params  # noqa: F401
from .synthetic import channel_obstacles, channel_params  # noqa: F401
