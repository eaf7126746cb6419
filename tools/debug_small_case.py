"""Debug helper: run a few steps of one small random case through the C ABI with a chosen kernel code and
report where the lattice differs from the oracle.  usage: python tools/debug_small_case.py [kernel [nx ny [steps]]]"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e.load_package(); orc = e.load_oracle()
from lbm_asynchronous_b200.lattice import make_param
nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (256, 20)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
p = orc.Params(nx, ny, 0, 10, 0.1, 0.005, 1.85)
rng = np.random.default_rng(0)
obst = (rng.random((ny, nx)) < 0.03).astype(np.int32)
cells0 = orc.init_cells(p); cells0 *= (1 + 0.05 * rng.standard_normal(cells0.shape)).astype(np.float32)
ref, _ = orc.run(p, obst, steps, cells=cells0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with pkg.Lattice(make_param(nx, ny, steps), obst, kernel=k, use_graph=False) as lat:
    lat.upload(cells0); lat.run(steps); c = lat.cells()
fl = obst == 0
bad = (c.view(np.uint32) != ref.view(np.uint32)).any(axis=2) & fl
rows = np.unique(np.nonzero(bad)[0])
print("kernel", k, f"{nx}x{ny} steps {steps} equal:", not bad.any(), "mismatching cells:", int(bad.sum()), "rows:", rows[:40], "..." if len(rows) > 40 else "")
if bad.any():
    for r in rows[:6]:
        cols = np.nonzero(bad[r])[0]
        print("  row", r, "cols", cols[:12], "..." if len(cols) > 12 else "", "count", len(cols))
