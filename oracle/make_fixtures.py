#!/usr/bin/env python3
"""Generates tests/golden/ from the reference -- TEST INFRASTRUCTURE, runs in the build container only.

Needs /root/reference (read-only mount) and oracle/_ref/ (`make -C oracle ref`: the reference's own
SerialCode and OpenMP programs compiled from the sources where they lie).  Nothing under
/root/reference is copied as source; what is committed is DATA:

  tests/golden/inputs/input_<grid>.params, obstacles_<grid>.dat
        the reference's four shipped problem instances (dataSet/), verbatim: they are the inputs of
        BASELINE.json configs 1-3.
  tests/golden/<grid>.npz, per shipped grid, arrays:
        golden_av_vels      float64[maxIters]  check/<grid>.av_vels.dat          (double-precision goldens)
        golden_pressure     float64[ny*nx]     check/<grid>.final_state.dat col 5 (128x128, 128x256 only)
        serial_av_vels      float32[maxIters]  av_vels.dat of the reference's SerialCode binary run here
                                               (1024x1024: the OpenMP binary; SerialCode needs ~9 min)
        serial_ux/uy/u/pressure float32[ny,nx] its final_state.dat (bit patterns: %.12E round-trips fp32)
                                               (1024x1024: pressure only, plus sha256 of the other planes)
        serial_program      which binary produced the serial_* arrays
        reynolds            the "Reynolds number" line of its stdout
  tests/golden/steps_<grid>.npz
        state after k = 1, 2, 3, 10, 101 steps (128x128 and 128x256), obtained by running the SerialCode
        BINARY on a params file with maxIters=k -- only av_vels/final_state are observable from the
        binary, so these hold (ux, uy, u, pressure) after k steps and av_vels[0:k].

Usage:  python oracle/make_fixtures.py [--runs /tmp/ref_runs]   (re-uses finished runs in --runs)
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
GRIDS = ["128x128", "128x256", "256x256", "1024x1024"]


def run_binary(prog: str, params: str, obstacles: str, outdir: str, threads: int | None = None) -> None:
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    with open(os.path.join(outdir, "stdout.txt"), "w") as fh:
        subprocess.run([os.path.join(HERE, "_ref", prog), params, obstacles], cwd=outdir, stdout=fh, stderr=subprocess.STDOUT,
                       check=True, env=env)


def load_final_state(path: str, nx: int, ny: int):
    a = np.loadtxt(path)
    assert a.shape == (nx * ny, 7)
    assert np.array_equal(a[:, 0].astype(int), np.tile(np.arange(nx), ny))
    assert np.array_equal(a[:, 1].astype(int), np.repeat(np.arange(ny), nx))
    planes = [a[:, c].astype(np.float32).reshape(ny, nx) for c in (2, 3, 4, 5)]
    for c, p in zip((2, 3, 4, 5), planes):  # 13 significant digits identify the fp32 value uniquely
        assert np.all(np.abs(p.astype(np.float64).ravel() - a[:, c]) <= 1e-12 * np.abs(a[:, c]))
    return planes


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", default="/tmp/ref_runs")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print("no /root/reference here: fixtures can only be generated in the build container", file=sys.stderr)
        return 1
    subprocess.run(["make", "-C", HERE, "all", "ref"], check=True)
    os.makedirs(os.path.join(GOLD, "inputs"), exist_ok=True)

    for g in GRIDS:
        pfile = os.path.join(REF, "dataSet", f"input_{g}.params")
        ofile = os.path.join(REF, "dataSet", f"obstacles_{g}.dat")
        shutil.copyfile(pfile, os.path.join(GOLD, "inputs", f"input_{g}.params"))
        shutil.copyfile(ofile, os.path.join(GOLD, "inputs", f"obstacles_{g}.dat"))
        tok = open(pfile).read().split()
        nx, ny = int(tok[0]), int(tok[1])

        prog = "d2q9-bgk-openmp" if g == "1024x1024" else "d2q9-bgk-serial"
        rdir = os.path.join(args.runs, ("openmp_" if g == "1024x1024" else "serial_") + g)
        if not os.path.exists(os.path.join(rdir, "final_state.dat")):
            print(f"running {prog} on {g} ...", flush=True)
            run_binary(prog, pfile, ofile, rdir)
        ux, uy, u, pr = load_final_state(os.path.join(rdir, "final_state.dat"), nx, ny)
        av = np.loadtxt(os.path.join(rdir, "av_vels.dat"), usecols=[1])
        av32 = av.astype(np.float32)
        assert np.all(np.abs(av32.astype(np.float64) - av) <= 1e-12 * np.abs(av))
        m = re.search(r"Reynolds number:\s+(\S+)", open(os.path.join(rdir, "stdout.txt")).read())
        out = dict(
            golden_av_vels=np.loadtxt(os.path.join(REF, "check", f"{g}.av_vels.dat"), usecols=[1]),
            serial_av_vels=av32,
            serial_pressure=pr,
            serial_program=np.array(prog),
            reynolds=np.float64(m.group(1)),
        )
        gfs = os.path.join(REF, "check", f"{g}.final_state.dat")
        if os.path.exists(gfs):
            out["golden_pressure"] = np.loadtxt(gfs, usecols=[5])
        if g == "1024x1024":
            for name, plane in (("ux", ux), ("uy", uy), ("u", u)):
                out[f"serial_{name}_sha256"] = np.array(hashlib.sha256(plane.tobytes()).hexdigest())
        else:
            out.update(serial_ux=ux, serial_uy=uy, serial_u=u)
        np.savez_compressed(os.path.join(GOLD, f"{g}.npz"), **out)
        print(f"{g}: wrote {g}.npz ({os.path.getsize(os.path.join(GOLD, g + '.npz')) / 1e6:.2f} MB)", flush=True)

    # short runs of the SerialCode binary: state after k steps (tests the first steps bit for bit)
    for g in ("128x128", "128x256"):
        pfile = os.path.join(REF, "dataSet", f"input_{g}.params")
        ofile = os.path.join(REF, "dataSet", f"obstacles_{g}.dat")
        tok = open(pfile).read().split()
        nx, ny = int(tok[0]), int(tok[1])
        out = {}
        with tempfile.TemporaryDirectory() as td:
            for k in (1, 2, 3, 10, 101):
                p2 = os.path.join(td, f"p{k}.params")
                with open(p2, "w") as fh:
                    fh.write("\n".join([tok[0], tok[1], str(k)] + tok[3:]) + "\n")
                rdir = os.path.join(td, f"run{k}")
                run_binary("d2q9-bgk-serial", p2, ofile, rdir)
                ux, uy, u, pr = load_final_state(os.path.join(rdir, "final_state.dat"), nx, ny)
                out[f"ux_{k}"], out[f"uy_{k}"], out[f"u_{k}"], out[f"pressure_{k}"] = ux, uy, u, pr
                out[f"av_vels_{k}"] = np.loadtxt(os.path.join(rdir, "av_vels.dat"), usecols=[1], ndmin=1).astype(np.float32)
        np.savez_compressed(os.path.join(GOLD, f"steps_{g}.npz"), **out)
        print(f"{g}: wrote steps_{g}.npz", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
