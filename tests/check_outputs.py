#!/usr/bin/env python3
"""`make check`: compare av_vels.dat / final_state.dat written by d2q9-bgk with a golden fixture, using
the rule of the reference's check/check.py (worst 100*(ref-sim)/sim must be finite and within the
tolerance, default 1 %; restated in oracle/pyoracle.py because the reference tree is not on the GPU box)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", required=True, help="tests/golden/<grid>.npz")
    ap.add_argument("--av-vels", required=True)
    ap.add_argument("--final-state", required=True)
    ap.add_argument("--tolerance", type=float, default=1.0)
    a = ap.parse_args()
    orc = entry.load_oracle()
    fx = np.load(a.golden)
    av = orc.read_av_vels(a.av_vels)
    fs = orc.read_final_state(a.final_state)
    ok = True
    for name, ref_av, ref_pr in (("shipped golden (double precision)", fx["golden_av_vels"], fx.get("golden_pressure")),
                                 ("SerialCode binary (fp32)", fx["serial_av_vels"], fx["serial_pressure"].ravel())):
        pa = orc.check_metric(ref_av, av)
        line = f"vs {name}: av_vels worst {pa:.3g} %"
        good = np.isfinite(pa) and abs(pa) <= a.tolerance
        if ref_pr is not None:
            pf = orc.check_metric(ref_pr, fs[:, 5])
            line += f", final_state pressure worst {pf:.3g} %"
            good = good and np.isfinite(pf) and abs(pf) <= a.tolerance
        print(line, "-> ok" if good else "-> FAILED")
        ok = ok and good
    print("Both tests passed!" if ok else "check failed")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
