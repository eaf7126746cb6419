"""step2_kernel: two timesteps per pass over HBM (lbm_fused2_kernel.cuh).  The strict flavour must give, after
every pair of steps, the bits the oracle (== SerialCode) gives after the same two steps -- on strips that
wrap in x, partial strips, segments that end inside a stage, boundary units that wrap in y or read a halo
ring, odd run lengths (a single-step kernel finishes the run) and runs split into several lbm_run calls."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_lattice_equal, bits, exact_tot_u, random_case, to_param

pytestmark = pytest.mark.gpu

F2 = 208432  # the default shape: 8 warps, stages of 4 rows, 3 stages, 2 CTAs per SM


@pytest.mark.parametrize("nx,ny,kernel,iters", [
    (128, 8, F2, 4), (128, 128, F2, 7), (132, 11, F2, 6), (240, 37, F2, 7), (248, 40, 216831, 8), (360, 9, 212441, 5),
    (1024, 70, 216831, 7), (2052, 23, 212441, 6), (4096, 300, 216831, 4), (640, 300, 212441, 9), (124 * 4, 64, 216831, 3),
    (128, 8, 212441, 4), (128, 128, 212441, 7), (2048, 600, 212441, 6), (2048, 600, 216831, 5),
])
def test_pairs_of_steps_bit_exact_vs_oracle(gpu, pkg, orc, nx, ny, kernel, iters):
    p, obst, cells0 = random_case(orc, nx, ny, seed=nx * 77 + ny)
    obst[:, 0] = (np.arange(ny) % 3 == 0)   # obstacles on both sides of the periodic seam in x
    obst[:, nx - 1] = (np.arange(ny) % 4 == 1)
    obst[ny - 2, :] = 0
    obst[ny - 2, :: max(1, nx // 7)] = 1
    ref_cells, ref_av = orc.run(p, obst, iters, cells=cells0)
    with pkg.Lattice(to_param(p, iters), obst, kernel=kernel) as lat:
        lat.upload(cells0)
        lat.run(iters)
        cells, av = lat.cells(), lat.av_vels()
        sums, bad = lat.tot_u_sums()
        fluid = lat.fluid_cells
        launches = lat.kernel_launches
    assert_lattice_equal(cells, ref_cells, obst)
    assert not bad.any()
    exact = exact_tot_u(orc, p, obst, cells0, iters)
    for t in range(iters):
        tot = (int(sums[t, 0]) + (int(sums[t, 1]) << 24)) * 2.0 ** -40
        assert abs(tot - exact[t][0]) <= fluid * 2.0 ** -41 + 1e-12 * exact[t][0]
    np.testing.assert_allclose(av, ref_av, rtol=5e-5)
    assert launches < 40 + iters  # pairs of steps: about half a launch per step, not two


@pytest.mark.parametrize("seg", [5, 6, 13, 31])
def test_segment_heights_that_end_inside_a_stage(gpu, pkg, orc, seg, monkeypatch):
    monkeypatch.setenv("LBM_F2_SEG", str(seg))
    p, obst, cells0 = random_case(orc, 384, 75, seed=seg, walls=False)  # no walls: the wrap in y carries flow
    ref_cells, _ = orc.run(p, obst, 6, cells=cells0)
    for kernel in (F2, 216831, 212441):
        with pkg.Lattice(to_param(p), obst, kernel=kernel) as lat:
            lat.upload(cells0)
            lat.run(6)
            assert_lattice_equal(lat.cells(), ref_cells, obst)


def test_runs_of_odd_and_even_length_chain(gpu, pkg, orc):
    """3 + 1 + 40 + 33 + 2 steps: pairs, a single step at the end of odd runs, graph replays, and accelerate-at-store
    never leaking across calls."""
    p, obst, cells0 = random_case(orc, 256, 48, seed=5)
    ref_cells, ref_av = orc.run(p, obst, 79, cells=cells0)
    with pkg.Lattice(to_param(p), obst, kernel=F2) as lat:
        lat.upload(cells0)
        avs = []
        for n in (3, 1, 40, 33, 2):
            lat.run(n)
            avs.append(lat.av_vels())
        assert lat.steps_done == 79
        assert_lattice_equal(lat.cells(), ref_cells, obst)
    np.testing.assert_allclose(np.concatenate(avs), ref_av, rtol=5e-5)


@pytest.mark.parametrize("kernel", [F2, 212441])
@pytest.mark.parametrize("nx,ny,n,runs", [(128, 64, 2, (6,)), (256, 90, 3, (5, 6, 33, 1)), (1024, 48, 4, (64, 3)), (132, 100, 5, (7, 2))])
def test_slabs_exchange_two_halo_rows_per_pair(gpu, pkg, orc, nx, ny, n, runs, kernel):
    """Several slabs (here on one device, step-major on one stream): the nine ring entries per side, the start-of-run
    push after accelerate_flow changed row ny-2, the neighbours' obstacle rows.  Same bits as the oracle, same
    integer |u| sums as one slab."""
    p, obst, cells0 = random_case(orc, nx, ny, seed=nx + ny + n, walls=False)
    obst[[0, ny - 1], ::3] = 1  # obstacles on the rows the neighbour recomputes
    obst[ny - 2, :] = 0
    total = sum(runs)
    ref_cells, _ = orc.run(p, obst, total, cells=cells0)
    with pkg.Lattice(to_param(p), obst, kernel=kernel) as lat:
        lat.upload(cells0)
        one = []
        for k in runs:
            lat.run(k)
            one.append(lat.tot_u_sums()[0])
    with pkg.Lattice(to_param(p), obst, devices=[0] * n, kernel=kernel) as lat:
        lat.upload(cells0)
        many = []
        for k in runs:
            lat.run(k)
            many.append(lat.tot_u_sums()[0])
        cells = lat.cells()
    assert_lattice_equal(cells, ref_cells, obst)
    for a, b in zip(one, many):
        assert np.array_equal(a[:, 0] + (a[:, 1] << 24), b[:, 0] + (b[:, 1] << 24))


def test_pairs_equal_single_steps_in_both_flavours(gpu, pkg, orc):
    """The same arithmetic in both kernels: pairs of steps == twice a single step, bit for bit, strict and fast."""
    p, obst, cells0 = random_case(orc, 512, 96, seed=11)
    for arith in ("strict", "fast"):
        outs = []
        for kernel in (F2, 11621):
            with pkg.Lattice(to_param(p), obst, kernel=kernel, arith=arith) as lat:
                lat.upload(cells0)
                lat.run(50)
                s = lat.tot_u_sums()[0]
                outs.append((lat.cells(), s[:, 0] + (s[:, 1] << 24)))
        if arith == "strict":
            assert np.array_equal(bits(outs[0][0]), bits(outs[1][0]))
            assert np.array_equal(outs[0][1], outs[1][1])
        else:  # the compiler may contract differently inside the two kernels
            np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-10)


def test_default_path_of_the_fast_flavour_on_a_large_grid_is_the_pair_kernel(gpu, pkg, orc):
    """kernel = 0, arith = fast on a grid too large for L2: about one launch per two steps; the strict flavour
    (HBM bound on single steps) keeps one interior + one boundary launch per step."""
    nx, ny, iters = 2048, 1024, 64
    obst = pkg.channel_obstacles(nx, ny)
    with pkg.Lattice(to_param(orc.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)), obst, arith="fast") as lat:
        l0 = lat.kernel_launches
        lat.run(iters)
        assert lat.kernel_launches - l0 <= iters // 2 + 8
    with pkg.Lattice(to_param(orc.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)), obst) as lat:
        l0 = lat.kernel_launches
        lat.run(iters)
        assert iters <= lat.kernel_launches - l0 <= 2 * iters + 8
