"""The host program's parallel obstacle-file reader (host/d2q9-bgk.c) against the reference's fscanf loop
(SerialCode/d2q9-bgk.c:588-601): same map, same four error messages, the first failing triple in file order
wins.  LBM_PARSE_ONLY=1 stops the program after parsing, so none of this needs a GPU.  Where oracle/_ref holds
the reference's own SerialCode binary, its stderr for the same malformed file is compared too."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import INPUTS, ROOT

EXE = os.path.join(ROOT, "lbm-asynchronous_b200", "d2q9-bgk")
REF = os.path.join(ROOT, "oracle", "_ref", "d2q9-bgk-serial")


def fnv1a_of_packed(obst: np.ndarray) -> int:
    rows, nx = obst.shape
    words = (nx + 31) // 32
    padded = np.zeros((rows, words * 32), dtype=np.uint8)
    padded[:, :nx] = obst != 0
    data = np.packbits(padded, axis=1, bitorder="little").tobytes()
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def parse_only(params, obstacles):
    return subprocess.run([EXE, str(params), str(obstacles)], capture_output=True, text=True, env=dict(os.environ, LBM_PARSE_ONLY="1"))


def write_params(path, nx, ny, iters=0):
    path.write_text(f"{nx}\n{ny}\n{iters}\n10\n0.1\n0.005\n1.85\n")


@pytest.mark.parametrize("grid", ["128x128", "128x256", "256x256", "1024x1024"])
def test_shipped_obstacle_files_parse_to_the_same_map(built, orc, grid):
    pf, of = os.path.join(INPUTS, f"input_{grid}.params"), os.path.join(INPUTS, f"obstacles_{grid}.dat")
    p = orc.read_params(pf)
    obst = orc.read_obstacles(of, p.nx, p.ny)
    r = parse_only(pf, of)
    assert r.returncode == 0, r.stderr
    m = re.search(r"blocked=(\d+) fnv1a=([0-9a-f]{16})", r.stdout)
    assert int(m.group(1)) == int((obst != 0).sum())
    assert int(m.group(2), 16) == fnv1a_of_packed(obst)


CASES = [
    ("1 2\n", "expected 3 values per line in obstacle file"),
    ("1 2 1\n3 4\n", "expected 3 values per line in obstacle file"),
    ("a b c\n", "expected 3 values per line in obstacle file"),
    ("1 2 x\n", "expected 3 values per line in obstacle file"),
    ("1 2 1 junk\n", "expected 3 values per line in obstacle file"),
    ("64 1 1\n", "obstacle x-coord out of range"),
    ("-1 1 1\n", "obstacle x-coord out of range"),
    ("1 32 1\n", "obstacle y-coord out of range"),
    ("1 1 0\n", "obstacle blocked value should be 1"),
    ("1 1 2\n", "obstacle blocked value should be 1"),
    ("1 1 1\n5 99 1\n99 1 1\n", "obstacle y-coord out of range"),   # the FIRST failing triple decides
    ("1 1 1\n99 99 7\n", "obstacle x-coord out of range"),           # checks in the reference's order
    ("1\n1\n1\n2 2 1", None),                                        # a triple may span lines; no trailing newline
    ("", None),
    ("  \n\n", None),
    ("+3 +4 +1\n", None),
]


@pytest.mark.parametrize("text,message", CASES)
def test_error_messages_are_the_references(built, tmp_path, text, message):
    pf, of = tmp_path / "p.params", tmp_path / "o.dat"
    write_params(pf, 64, 32)
    of.write_text(text)
    r = parse_only(pf, of)
    if message is None:
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1
        lines = r.stderr.splitlines()
        assert lines[0].startswith("Error at line ") and lines[1] == message
    if os.path.exists(REF):  # the reference's own program on the same file
        ref = subprocess.run([REF, str(pf), str(of)], cwd=tmp_path, capture_output=True, text=True)
        if message is None:
            assert ref.returncode == 0
        else:
            assert ref.returncode == 1 and ref.stderr.splitlines()[1] == message


def test_large_file_is_split_between_threads_and_errors_keep_file_order(built, tmp_path):
    """~3 MB of valid lines (several chunks), then the same with a bad line in the middle and a different bad
    line near the end: the earlier one is reported."""
    nx, ny = 4096, 4096
    rng = np.random.default_rng(1)
    xs, ys = rng.integers(0, nx, 300000), rng.integers(0, ny, 300000)
    lines = [f"{x} {y} 1" for x, y in zip(xs, ys)]
    pf, of = tmp_path / "p.params", tmp_path / "o.dat"
    write_params(pf, nx, ny)
    of.write_text("\n".join(lines) + "\n")
    r = parse_only(pf, of)
    assert r.returncode == 0, r.stderr
    obst = np.zeros((ny, nx), dtype=np.int32)
    obst[ys, xs] = 1
    m = re.search(r"blocked=(\d+) fnv1a=([0-9a-f]{16})", r.stdout)
    assert int(m.group(1)) == int(obst.sum()) and int(m.group(2), 16) == fnv1a_of_packed(obst)
    bad = list(lines)
    bad[150000] = "5 5 0"
    bad[290000] = "99999 1 1"
    of.write_text("\n".join(bad) + "\n")
    r = parse_only(pf, of)
    assert r.returncode == 1 and r.stderr.splitlines()[1] == "obstacle blocked value should be 1"
    bad[100] = "7 zz 1"  # a non-integer earlier than both: everything after it is never converted
    of.write_text("\n".join(bad) + "\n")
    r = parse_only(pf, of)
    assert r.returncode == 1 and r.stderr.splitlines()[1] == "expected 3 values per line in obstacle file"
