"""tests/ref_check: the verbatim check.py and the golden .dat writer (no GPU)."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

REFCHECK = os.path.join(ROOT, "tests", "ref_check")
REF = "/root/reference/check"
CHECK_PY_SHA256 = "44673f5a20d0790fc32940d198964b94fcc3805194fb53a397f8ebc55a92f8b0"


def test_check_py_is_the_references_file_unchanged():
    with open(os.path.join(REFCHECK, "check.py"), "rb") as fh:
        mine = fh.read()
    assert hashlib.sha256(mine).hexdigest() == CHECK_PY_SHA256
    if os.path.exists(os.path.join(REF, "check.py")):  # build container only
        with open(os.path.join(REF, "check.py"), "rb") as fh:
            assert fh.read() == mine


def run_check(ref_av, ref_fs, av, fs):
    return subprocess.run([sys.executable, os.path.join(REFCHECK, "check.py"), f"--ref-av-vels-file={ref_av}",
                           f"--ref-final-state-file={ref_fs}", f"--av-vels-file={av}", f"--final-state-file={fs}"],
                          capture_output=True, text=True)


def test_written_goldens_feed_check_py_and_match_the_shipped_files(tmp_path, orc):
    r = subprocess.run([sys.executable, os.path.join(REFCHECK, "write_goldens.py"), str(tmp_path), "128x128"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    av, fs = tmp_path / "128x128.av_vels.dat", tmp_path / "128x128.final_state.dat"
    if os.path.exists(os.path.join(REF, "128x128.av_vels.dat")):
        with open(av, "rb") as a, open(os.path.join(REF, "128x128.av_vels.dat"), "rb") as b:
            assert a.read() == b.read()
        mine = np.loadtxt(fs, usecols=[0, 1, 5])
        shipped = np.loadtxt(os.path.join(REF, "128x128.final_state.dat"), usecols=[0, 1, 5])
        assert np.array_equal(mine, shipped)
    # the SerialCode fixture (fp32 run of the reference binary) passes the real check.py against the goldens ...
    fx = np.load(os.path.join(GOLDEN, "128x128.npz"))
    sim_av, sim_fs = tmp_path / "sim_av.dat", tmp_path / "sim_fs.dat"
    with open(sim_av, "w") as fh:
        fh.write("".join(f"{t}:\t{v:.12E}\n" for t, v in enumerate(fx["serial_av_vels"])))
    ny, nx = fx["serial_pressure"].shape
    with open(sim_fs, "w") as fh:
        for jj in range(ny):
            fh.write("".join(f"{ii} {jj} {fx['serial_ux'][jj, ii]:.12E} {fx['serial_uy'][jj, ii]:.12E} {fx['serial_u'][jj, ii]:.12E} "
                             f"{fx['serial_pressure'][jj, ii]:.12E} 0\n" for ii in range(nx)))
    c = run_check(av, fs, sim_av, sim_fs)
    assert c.returncode == 0 and "Both tests passed!" in c.stdout, c.stdout + c.stderr
    # ... and agrees with the oracle's restatement of its metric (pyoracle.check_passes)
    ok, a, f = orc.check_passes(fx["golden_av_vels"], fx["serial_av_vels"], fx["golden_pressure"], fx["serial_pressure"].ravel())
    assert ok
    # a 2 % perturbation of av_vels fails it
    with open(sim_av, "w") as fh:
        fh.write("".join(f"{t}:\t{v * 1.02:.12E}\n" for t, v in enumerate(fx["serial_av_vels"])))
    c = run_check(av, fs, sim_av, sim_fs)
    assert c.returncode == 1 and "av_vels failed check" in c.stdout
