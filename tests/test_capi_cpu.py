"""CPU-side checks of the C-ABI library and the host program: loads, exports, host logic, error
behaviour.  No compute call needs a GPU here."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import INPUTS, ROOT, has_gpu

PKG_DIR = os.path.join(ROOT, "lbm-asynchronous_b200")


def test_library_exports_every_symbol_the_header_declares(pkg):
    from lbm_asynchronous_b200 import capi

    header = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS)
    lib = pkg.library()
    for name in declared:
        assert getattr(lib, name) is not None
    nm = subprocess.run(["nm", "-D", "--defined-only", capi.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lbm_[a-z0-9_]+)", nm))
    assert declared <= exported


def test_param_struct_layout_is_t_param(pkg):
    import ctypes as C

    from lbm_asynchronous_b200.capi import Param

    # SerialCode/d2q9-bgk.c:66-75: four ints then three floats
    assert C.sizeof(Param) == 28
    assert [f[0] for f in Param._fields_] == ["nx", "ny", "maxIters", "reynolds_dim", "density", "accel", "omega"]


def test_partition_is_balanced_and_keeps_driven_row_interior(pkg):
    for ny, n in [(128, 2), (128, 3), (1024, 8), (1024, 4), (32768, 8), (11, 4), (7, 2)]:
        st = pkg.partition(ny, n)
        assert st[0] == 0 and st[-1] == ny and len(st) == n + 1
        rows = np.diff(st)
        assert rows.max() - rows.min() <= 1
        assert all(r >= 2 for r in rows[:-1]) and rows[-1] >= 3
        assert st[-2] < ny - 2 < ny - 1  # the driven row ny-2 is interior to the last slab
    assert pkg.partition(128, 1) == [0, 128]
    with pytest.raises(pkg.LbmError) as e:
        pkg.partition(8, 4)
    assert e.value.code == 1 and "too small" in str(e.value)


def test_av_from_sums_is_the_reference_division(pkg):
    # total = lo + hi * 2^24 in units of 2^-40; tot_u rounded to float, then / (float)cells
    one = 1 << 40
    assert pkg.av_from_sums(one & 0xFFFFFF, one >> 24, 0, 4) == np.float32(0.25)
    lo, hi = 123456, 7_000_000
    tot = np.float32((lo + (hi << 24)) * 2.0 ** -40)
    assert pkg.av_from_sums(lo, hi, 0, 15876) == tot / np.float32(15876)
    assert np.isnan(pkg.av_from_sums(1, 1, 3, 10))


@pytest.mark.skipif(has_gpu(), reason="checks the behaviour on a box without a CUDA device")
def test_no_cpu_fallback_compute_fails_loudly_without_a_device(pkg):
    from lbm_asynchronous_b200.lattice import make_param

    assert pkg.library().lbm_device_count() == 0
    with pytest.raises(pkg.LbmError) as e:
        pkg.Lattice(make_param(16, 16, 1), np.zeros((16, 16), np.int32))
    assert e.value.code == 2 and "no CUDA device" in str(e.value)
    with pytest.raises(pkg.LbmError) as e:
        pkg.SlabLattice(make_param(16, 16, 1), np.zeros((16, 16), np.int32), 0, 16, 0, 1, 0)
    assert e.value.code == 2


def test_product_never_touches_the_oracle():
    """Nothing under the package directory, include/ or the Makefile's product targets may reference oracle/."""
    for base, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                clean = ("pyoracle" not in text) and ("lbm_oracle" not in text) and ("oracle/" not in text)
                assert clean, f
    assert "oracle" not in open(os.path.join(ROOT, "include", "lbm_b200.h")).read()


# ---- the host program: CLI and error behaviour of the reference (SerialCode/d2q9-bgk.c:745-757) ----
BIN = os.path.join(PKG_DIR, "d2q9-bgk")


def run_host(args, cwd):
    return subprocess.run([BIN] + args, cwd=cwd, capture_output=True, text=True)


def test_usage_message_and_exit_status(built, tmp_path):
    r = run_host([], tmp_path)
    assert r.returncode == 1
    assert r.stderr == f"Usage: {BIN} <paramfile> <obstaclefile>\n"
    r = run_host(["a", "b", "c"], tmp_path)
    assert r.returncode == 1 and r.stderr.startswith("Usage:")


def test_param_file_errors(built, tmp_path):
    r = run_host(["/nonexistent.params", "x"], tmp_path)
    assert r.returncode == 1
    assert "could not open input parameter file: /nonexistent.params" in r.stderr
    assert re.match(r"Error at line \d+ of file .*d2q9-bgk\.c:\n", r.stderr)
    names = ["nx", "ny", "maxIters", "reynolds_dim", "density", "accel", "omega"]
    good = ["128", "128", "10", "10", "0.1", "0.005", "1.85"]
    for i, name in enumerate(names):
        pf = tmp_path / f"bad_{name}.params"
        pf.write_text("\n".join(good[:i] + ["oops"] + good[i + 1:]) + "\n")
        r = run_host([str(pf), "x"], tmp_path)
        assert r.returncode == 1 and f"could not read param file: {name}\n" in r.stderr, (name, r.stderr)


def test_obstacle_file_errors(built, tmp_path):
    pf = os.path.join(INPUTS, "input_128x128.params")
    r = run_host([pf, "/nonexistent.dat"], tmp_path)
    assert r.returncode == 1 and "could not open input obstacles file: /nonexistent.dat" in r.stderr
    cases = {
        "1 2\n": "expected 3 values per line in obstacle file",
        "128 0 1\n": "obstacle x-coord out of range",
        "-1 0 1\n": "obstacle x-coord out of range",
        "0 128 1\n": "obstacle y-coord out of range",
        "0 0 2\n": "obstacle blocked value should be 1",
    }
    for i, (text, msg) in enumerate(cases.items()):
        of = tmp_path / f"o{i}.dat"
        of.write_text("3 3 1\n" + text)
        r = run_host([pf, str(of)], tmp_path)
        assert r.returncode == 1 and msg in r.stderr, (text, r.stderr)


@pytest.mark.skipif(has_gpu(), reason="checks the behaviour on a box without a CUDA device")
def test_host_program_dies_without_a_device(built, tmp_path):
    r = run_host([os.path.join(INPUTS, "input_128x128.params"), os.path.join(INPUTS, "obstacles_128x128.dat")], tmp_path)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "Error at line" in r.stderr
    assert not os.path.exists(tmp_path / "final_state.dat")


def test_gen_channel_matches_the_numpy_generator(built, pkg, orc, tmp_path):
    nx, ny = 64, 40
    r = subprocess.run([os.path.join(PKG_DIR, "gen_channel"), str(nx), str(ny), "7", "p.params", "o.dat", "0.05", "42"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p = orc.read_params(str(tmp_path / "p.params"))
    assert (p.nx, p.ny, p.max_iters, p.reynolds_dim) == (nx, ny, 7, 10)
    assert (p.density, p.accel, p.omega) == (np.float32(0.1), np.float32(0.005), np.float32(1.85))
    obst = orc.read_obstacles(str(tmp_path / "o.dat"), nx, ny)
    want = pkg.channel_obstacles(nx, ny, p=0.05, seed=42)
    assert np.array_equal(obst, want)
    assert want[0].all() and want[-1].all() and not want[ny - 2].any()
    assert 0.02 < want[1:ny - 2].mean() < 0.08
    # a slab of rows is the same as slicing the full map
    assert np.array_equal(pkg.channel_obstacles(nx, ny, p=0.05, seed=42, row0=13, row1=31), want[13:31])


def test_missing_library_fails_loudly(pkg, monkeypatch):
    """No silent fallback: if liblbm_b200.so is not there, the Python layer raises instead of computing
    anything some other way."""
    from lbm_asynchronous_b200 import capi

    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_LIB", "/nonexistent/liblbm_b200.so")
    with pytest.raises(capi.LbmError) as e:
        capi.library()
    assert "is missing" in str(e.value) and "no CPU or PyTorch fallback" in str(e.value)
