"""Row slabs on SEVERAL B200s: peer stores into the neighbour GPU's halo ring over NVLink.
Skipped (not failed) on a box with a single GPU; the single-GPU tests already exercise the halo
protocol with several slabs on one device."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS, ROOT

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def load_case(orc, grid):
    p = orc.read_params(os.path.join(INPUTS, f"input_{grid}.params"))
    obst = orc.read_obstacles(os.path.join(INPUTS, f"obstacles_{grid}.dat"), p.nx, p.ny)
    return p, obst


def to_param(p, iters=None):
    from lbm_asynchronous_b200.lattice import make_param

    return make_param(p.nx, p.ny, p.max_iters if iters is None else iters, p.reynolds_dim, p.density, p.accel, p.omega)


@pytest.fixture()
def ngpu(gpu):
    if gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    return gpu


@pytest.mark.parametrize("grid,iters", [("1024x1024", 1500), ("128x256", 3000)])
def test_sync_slabs_on_n_gpus_equal_one_gpu(ngpu, pkg, orc, grid, iters):
    """BASELINE config 3: the N-GPU result with synchronous halos equals the 1-GPU result bit for bit
    (lattice AND the integer |u| sums)."""
    p, obst = load_case(orc, grid)
    with pkg.Lattice(to_param(p), obst, ngpus=1) as lat:
        lat.run(iters)
        one = (lat.cells(), lat.tot_u_sums()[0])
    for n in sorted({2, min(4, ngpu), ngpu}):
        with pkg.Lattice(to_param(p), obst, ngpus=n) as lat:
            assert [s[2] for s in lat.slabs()] == list(range(n))
            lat.run(iters)
            many = (lat.cells(), lat.tot_u_sums()[0])
        assert np.array_equal(bits(one[0]), bits(many[0])), n
        assert np.array_equal(one[1][:, 0] + (one[1][:, 1] << 24), many[1][:, 0] + (many[1][:, 1] << 24)), n


def test_resident_step_loop_across_gpus(ngpu, pkg, orc):
    """kernel=200: step_loop_kernel on every GPU at once, halo rows and flags exchanged between the resident
    kernels (not the default across GPUs: it measured no faster than the graph path).  Same bits."""
    p, obst = load_case(orc, "1024x1024")
    iters = 600
    with pkg.Lattice(to_param(p), obst, ngpus=1) as lat:
        lat.run(iters)
        one = (lat.cells(), lat.tot_u_sums()[0])
    for n, kernel in ((2, 200), (min(4, ngpu), 204), (2, 201)):
        with pkg.Lattice(to_param(p), obst, ngpus=n, kernel=kernel) as lat:
            lat.run(250)
            lat.run(350)
            many = (lat.cells(), lat.tot_u_sums()[0])
            launches = lat.kernel_launches
        assert np.array_equal(bits(one[0]), bits(many[0])), (n, kernel)
        assert launches < 100 * n  # set-up kernels plus a handful of cooperative launches, not one per step


@pytest.mark.parametrize("kernel", [401, 402, 404, 514])
def test_packet_kernel_across_gpus(ngpu, pkg, orc, kernel):
    """step_ll_kernel (401, 402, 404: one, two, four cells per thread) / step_band_kernel (514) on every GPU at once: rows inside a slab exchange packets through L2
    (or bands their rows), the slabs' boundary
    rows through each other's memory over NVLink (st / ld.relaxed.sys.b128), every run seeded with the packets of
    the current state.  Same bits as one GPU, also over several lbm_run calls, on a shipped case whose periodic wrap
    crosses the GPU 0 <-> GPU N-1 link and on a random state."""
    p, obst = load_case(orc, "128x256")
    with pkg.Lattice(to_param(p), obst, ngpus=1, kernel=201) as lat:
        lat.run(600)
        one = (lat.cells(), lat.tot_u_sums()[0])
    for n in sorted({2, min(4, ngpu)}):
        with pkg.Lattice(to_param(p), obst, ngpus=n, kernel=kernel) as lat:
            sums = []
            for it in (1, 249, 350):
                lat.run(it)
                sums.append(lat.tot_u_sums()[0])
            many = lat.cells()
            launches = lat.kernel_launches
        tot = np.concatenate(sums)
        assert np.array_equal(bits(one[0]), bits(many)), (n, kernel)
        assert np.array_equal(one[1][:, 0] + (one[1][:, 1] << 24), tot[:, 0] + (tot[:, 1] << 24)), (n, kernel)
        assert launches < 300  # one cooperative launch (+ seed) per slab and run after the set-up kernels, not one per step
    from test_gpu_parity import random_case

    p, obst, cells0 = random_case(orc, 256, 96, seed=12)
    ref_cells, ref_av = orc.run(p, obst, 9, cells=cells0)
    with pkg.Lattice(to_param(p, 9), obst, ngpus=2, kernel=kernel) as lat:
        lat.upload(cells0)
        lat.run(4)
        lat.run(5)
        cells = lat.cells()
    fluid = obst == 0
    assert np.array_equal(bits(cells[fluid]), bits(ref_cells[fluid]))


def test_packet_kernel_one_process_per_gpu(ngpu, built, tmp_path):
    """The same through the per-process front end (torchrun, packet areas mapped through CUDA IPC)."""
    out = tmp_path / "result.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "sharded_worker.py"), str(out), "128x256", "402"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = np.load(out)
    assert bool(res["cells_equal"]) and bool(res["av_equal"])


def test_async_halo_mode_drift_is_small(ngpu, pkg, orc):
    """The stale-halo mode (un-waited MPI_Testall): free running, so no bit-exact expectation and no
    reference fixture (parity unpinned).  Its drift against the synchronous run is
    REPORTED with check.py's metric (-s shows it); the assertion only bounds it loosely: the
    staleness is unbounded by construction, like the reference's."""
    p, obst = load_case(orc, "1024x1024")
    iters = 2000
    with pkg.Lattice(to_param(p), obst, ngpus=1) as lat:
        lat.run(iters)
        av_ref = lat.av_vels()
        pr_ref = lat.final_state()[3]
    with pkg.Lattice(to_param(p), obst, ngpus=min(ngpu, 4), halo_mode="async") as lat:
        lat.run(iters)
        av = lat.av_vels()
        pr = lat.final_state()[3]
    a = orc.check_metric(av_ref, av)
    f = orc.check_metric(pr_ref.ravel(), pr.ravel())
    print(f"async drift vs sync after {iters} steps: av_vels {a:.3g} %, pressure {f:.3g} %")
    assert np.isfinite(a) and np.isfinite(f) and abs(a) < 10.0 and abs(f) < 10.0


def test_host_program_on_two_gpus(ngpu, built, orc, tmp_path):
    exe = os.path.join(ROOT, "lbm-asynchronous_b200", "d2q9-bgk")
    grid = "128x256"  # open top/bottom rows: the periodic wrap crosses the GPU 0 <-> GPU N-1 link
    env = dict(os.environ, LBM_GPUS="2")
    r = subprocess.run([exe, os.path.join(INPUTS, f"input_{grid}.params"), os.path.join(INPUTS, f"obstacles_{grid}.dat")],
                       cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "gpus=2 slabs=2" in r.stdout
    fx = np.load(os.path.join(GOLDEN, f"{grid}.npz"))
    av = orc.read_av_vels(str(tmp_path / "av_vels.dat"))
    fs = orc.read_final_state(str(tmp_path / "final_state.dat"))
    ok, a, f = orc.check_passes(fx["golden_av_vels"], av, fx["golden_pressure"], fs[:, 5])
    assert ok, (a, f)
    assert np.array_equal(fs[:, 5].astype(np.float32).view(np.uint32), bits(fx["serial_pressure"]).ravel())


def test_one_process_per_gpu_under_torchrun(ngpu, built, tmp_path):
    """The bench's topology: torchrun, one rank per GPU, rings mapped through CUDA IPC."""
    n = 2
    out = tmp_path / "result.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "sharded_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = np.load(out)
    assert bool(res["cells_equal"]) and bool(res["av_equal"])


@pytest.mark.parametrize("lag", [0, 2])
def test_deterministic_stale_halo_on_real_gpus(ngpu, pkg, orc, lag):
    """halo_lag on slabs that live on different GPUs (graph path: TMA interior kernel + boundary kernel with
    flags over NVLink) against the oracle's decomposed run."""
    p, obst = load_case(orc, "128x256")
    iters = 200
    n = 2
    starts = pkg.partition(p.ny, n)
    ref_cells, ref_av = orc.run_decomposed(p, obst, starts, lag, iters)
    with pkg.Lattice(to_param(p), obst, ngpus=n, halo_lag=lag) as lat:
        lat.run(iters)
        cells, av = lat.cells(), lat.av_vels()
    fluid = obst == 0
    assert np.array_equal(bits(cells[fluid]), bits(ref_cells[fluid]))
    np.testing.assert_allclose(av, ref_av, rtol=5e-5)
