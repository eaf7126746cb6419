"""Debug helper: run a few steps of one small random case through the C ABI with a chosen kernel code and
report where the lattice differs from the oracle.  usage: python tools/debug_small_case.py [kernel [nx ny]]"""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as e
pkg = e.load_package(); orc = e.load_oracle()
from lbm_asynchronous_b200.lattice import make_param
nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (256, 20)
p = orc.Params(nx, ny, 0, 10, 0.1, 0.005, 1.85)
rng = np.random.default_rng(0)
obst = (rng.random((ny, nx)) < 0.03).astype(np.int32)
cells0 = orc.init_cells(p); cells0 *= (1 + 0.05 * rng.standard_normal(cells0.shape)).astype(np.float32)
ref, _ = orc.run(p, obst, 3, cells=cells0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with pkg.Lattice(make_param(nx, ny, 3), obst, kernel=k, use_graph=False) as lat:
    lat.upload(cells0); lat.run(3); c = lat.cells()
fl = obst == 0
print("kernel", k, "equal:", np.array_equal(c[fl].view(np.uint32), ref[fl].view(np.uint32)), "mismatch rows:", np.unique(np.nonzero((c.view(np.uint32) != ref.view(np.uint32)).any(axis=2) & fl)[0])[:20])
