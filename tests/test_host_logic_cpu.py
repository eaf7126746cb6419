"""Property tests (hypothesis) of the host-side logic behind the C ABI -- no GPU needed."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st


@settings(max_examples=300, deadline=None)
@given(ny=st.integers(3, 100000), n=st.integers(1, 64))
def test_partition_properties(pkg, ny, n):
    """lbm_partition: slabs tile [0, ny) in order, sizes differ by at most one, every slab has >= 2 rows and
    the last >= 3 (so the driven row ny-2 is interior to it) -- or the call is refused as too small."""
    try:
        s = pkg.partition(ny, n)
    except pkg.LbmError as e:
        assert e.code == 1
        base, rem = divmod(ny, n)
        smallest_last = base + (1 if rem else 0)
        assert n > 1 and (base < 2 or smallest_last < 3)
        return
    assert s[0] == 0 and s[-1] == ny and len(s) == n + 1
    rows = np.diff(s)
    assert (rows > 0).all() and rows.max() - rows.min() <= 1
    if n > 1:
        assert rows[:-1].min() >= 2 and rows[-1] >= 3 and s[-2] < ny - 2


@settings(max_examples=300, deadline=None)
@given(lo=st.integers(0, (1 << 62) - 1), hi=st.integers(0, (1 << 40) - 1), cells=st.integers(1, 1 << 31))
def test_av_from_sums_is_exact_integer_total_then_reference_division(pkg, lo, hi, cells):
    """tot_u = float((lo + hi*2^24) * 2^-40) with one rounding from the exact integer, then the reference's
    fp32 division tot_u / (float)tot_cells (SerialCode/d2q9-bgk.c:457)."""
    total = lo + (hi << 24)
    # the library rounds the exact 128-bit integer to double first (hi64*2^64 + lo64 in double), then to float
    hi64, lo64 = total >> 64, total & ((1 << 64) - 1)
    as_double = (float(hi64) * 18446744073709551616.0 + float(lo64)) * (1.0 / 1099511627776.0)
    want = np.float32(as_double) / np.float32(cells)
    got = pkg.av_from_sums(lo, hi, 0, cells)
    assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32)
    # and it is within one float ulp of the true quotient
    true = total / 2.0 ** 40 / cells
    if true > 1e-30:
        assert abs(float(got) - true) <= 2.0 ** -22 * true


@settings(max_examples=25, deadline=None)
@given(nx=st.integers(1, 300), ny=st.integers(4, 200), a=st.integers(0, 199), b=st.integers(0, 199), seed=st.integers(0, 2 ** 32))
def test_channel_generator_slabs_are_slices(pkg, nx, ny, a, b, seed):
    r0, r1 = sorted((a % ny, b % ny))
    r1 += 1
    full = pkg.channel_obstacles(nx, ny, p=0.1, seed=seed)
    assert np.array_equal(pkg.channel_obstacles(nx, ny, p=0.1, seed=seed, row0=r0, row1=r1), full[r0:r1])
    assert full[0].all() and full[-1].all() and not full[ny - 2].any()
    assert set(np.unique(full)) <= {0, 1}
