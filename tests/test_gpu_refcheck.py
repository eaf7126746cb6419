"""The reference's OWN acceptance gate (check/check.py, byte for byte under tests/ref_check/) run on the
files the `d2q9-bgk` host program writes, for all four shipped grids at their full iteration counts
(SerialCode/Makefile:20-25 wires the same script to the same file names).  BASELINE.json: "check.py
passes on all four shipped grids"."""
import os
import re
import subprocess
import sys

import pytest

from conftest import INPUTS, ROOT

pytestmark = pytest.mark.gpu

REFCHECK = os.path.join(ROOT, "tests", "ref_check")


@pytest.fixture(scope="module")
def golden_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("ref_goldens")
    r = subprocess.run([sys.executable, os.path.join(REFCHECK, "write_goldens.py"), str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return d


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("grid", ["128x128", "128x256", "256x256", "1024x1024"])
def test_reference_check_py_passes_on_cli_output(gpu, built, golden_dir, tmp_path, grid, arith):
    exe = os.path.join(ROOT, "lbm-asynchronous_b200", "d2q9-bgk")
    env = dict(os.environ, LBM_ARITH=arith)
    r = subprocess.run([exe, os.path.join(INPUTS, f"input_{grid}.params"), os.path.join(INPUTS, f"obstacles_{grid}.dat")],
                       cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    c = subprocess.run([sys.executable, os.path.join(REFCHECK, "check.py"),
                        f"--ref-av-vels-file={golden_dir}/{grid}.av_vels.dat",
                        f"--ref-final-state-file={golden_dir}/{grid}.final_state.dat",
                        f"--av-vels-file={tmp_path}/av_vels.dat", f"--final-state-file={tmp_path}/final_state.dat"],
                       capture_output=True, text=True)
    pcts = re.findall(r"= (\S+)%", c.stdout)
    print(f"check.py {grid} {arith}: av_vels {pcts[0] if pcts else '?'} %, final_state {pcts[1] if len(pcts) > 1 else '?'} %")
    assert c.returncode == 0 and "Both tests passed!" in c.stdout, c.stdout + c.stderr
