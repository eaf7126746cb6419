"""Parity of the CUDA path (through the C ABI) with the oracle and with the reference's own outputs.

Bars: the STRICT flavour is bit-exact on the lattice (integer compare of the fp32 bit patterns) --
against the oracle on seeded inputs and against the SerialCode binary's final_state after the FULL
40 000 / 80 000 / 20 000 steps of the shipped cases; av_vels is an exact integer reduction, compared
with the oracle's double-precision sum of the same per-cell values (tolerance: 2^-41 per cell from
the fixed-point rounding) and with the reference's sequential fp32 sum at rtol 5e-5 (summation
order).  The FAST flavour is held to check.py's rule (1 %) against goldens and SerialCode, and to
5e-3 % (pressure) on the full runs.
"""
import os
import subprocess

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


def random_case(orc, nx, ny, seed, p_obst=0.03, walls=True):
    rng = np.random.default_rng(seed)
    p = orc.Params(nx, ny, 0, 10, 0.1, 0.005, 1.85)
    obst = (rng.random((ny, nx)) < p_obst).astype(np.int32)
    if walls:
        obst[0, :] = 1
        obst[-1, :] = 1
    if ny >= 2:
        obst[ny - 2, :: max(1, nx // 5)] = 1  # a few blocked cells on the driven row too
    cells = orc.init_cells(p)
    cells *= (1 + 0.05 * rng.standard_normal(cells.shape)).astype(np.float32)
    return p, obst, cells


def assert_lattice_equal(cells, ref_cells, obst):
    fluid = obst == 0
    assert np.array_equal(bits(cells[fluid]), bits(ref_cells[fluid]))
    # obstacle cells: speeds 1..8 hold the bounce-back values (speed 0 is a don't-care of the reference)
    assert np.array_equal(bits(cells[~fluid][:, 1:]), bits(ref_cells[~fluid][:, 1:]))


def exact_tot_u(orc, p, obst, cells_before, k):
    """double-precision sum of the per-cell fp32 |u| after each of k steps (oracle)"""
    out = []
    c = cells_before
    for _ in range(k):
        c, _ = orc.run(p, obst, 1, cells=c)
        out.append(orc.tot_u_f64(p, c, obst))
    return out


# ---------------------------------------------------------------------------------------------
# strict flavour vs the oracle, seeded inputs, every kernel shape
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,kernel,block", [
    (128, 128, 0, 0), (128, 64, 10, 128), (256, 37, 21, 256), (512, 16, 32, 512), (1024, 9, 0, 256),
    (64, 50, 0, 0), (32, 33, 0, 0), (8, 12, 0, 0), (4, 5, 0, 0),          # warps that span several rows
    (100, 30, 0, 0), (36, 21, 0, 128), (2052, 7, 0, 0),                    # nx % 4 == 0 but not a power of two
    (127, 20, 0, 0), (33, 17, 0, 128), (1, 8, 0, 0), (3, 3, 0, 0), (130, 2, 0, 0),  # scalar kernel
    (128, 40, 99, 0),                                                       # scalar kernel forced
    # step_tma_kernel for the interior rows (TY rows per tile, stages): full, partial and single tiles
    (128, 128, 10832, 0), (128, 3, 10832, 0), (128, 4, 10823, 0), (256, 37, 10841, 0), (384, 20, 10443, 0),
    (132, 11, 10462, 0), (1024, 9, 11621, 0), (640, 40, 11631, 0), (2052, 7, 10822, 0), (4096, 70, 10434, 128),
    (256, 45, 11041, 0), (1024, 23, 11031, 0), (640, 300, 11231, 0), (2052, 130, 11221, 0), (4096, 333, 11041, 0),
    (128, 64, 32, 256),                                                     # step_vec4_kernel for every row
    (256, 37, 200, 128), (1024, 700, 200, 0), (4096, 200, 204, 256), (12, 9, 200, 0),  # step_loop_kernel (cooperative)
    (127, 33, 201, 0), (640, 300, 201, 0), (1, 5, 200, 0), (64, 64, 204, 0),
    # step_cluster_kernel (lattice resident in one cluster's shared memory): 3000 = shape chosen by the library,
    # 3CVM = C cells per thread, V both pairs side by side, M x 256 threads; ragged row blocks, CTAs without rows,
    # fewer rows than CTAs, rows shorter than a warp
    (128, 128, 3000, 0), (128, 128, 3401, 0), (128, 128, 3202, 0), (128, 128, 3104, 0), (128, 64, 3411, 0),
    (256, 37, 3402, 0), (64, 50, 3202, 0), (127, 20, 3104, 0), (100, 30, 3000, 0), (32, 33, 3000, 0), (8, 12, 3000, 0),
    (4, 5, 3000, 0), (1, 8, 3000, 0), (3, 3, 3000, 0), (130, 2, 3000, 0), (36, 21, 3401, 0), (128, 256, 3000, 0),
    (256, 128, 3404, 0), (256, 128, 3204, 0), (512, 16, 3000, 0), (1024, 7, 3000, 0),
    # step_ll_kernel (a row per CTA, cells in registers, flagged packets between rows): partial warps, one-row and
    # one-column grids, the widest row a CTA takes
    (128, 128, 400, 0), (128, 256, 400, 0), (256, 256, 400, 0), (64, 50, 400, 0), (127, 20, 400, 0), (100, 30, 400, 0),
    (32, 33, 400, 0), (8, 12, 400, 0), (4, 5, 400, 0), (1, 8, 400, 0), (3, 3, 400, 0), (130, 2, 400, 0), (5, 2, 400, 0),
    (1024, 9, 400, 0), (1000, 2, 400, 0), (333, 290, 400, 0), (128, 64, 401, 0),
    # the same with four cells per thread (packed collision, four packets per direction)
    (128, 128, 404, 0), (1024, 256, 404, 0), (256, 37, 404, 0), (8, 12, 404, 0), (4, 5, 404, 0), (100, 30, 404, 0),
    (36, 21, 404, 0), (512, 280, 404, 0), (1024, 2, 404, 0),
    (1024, 256, 402, 0), (128, 128, 402, 0), (100, 30, 402, 0), (6, 5, 402, 0), (258, 40, 402, 0),
    # step_band_kernel (a band of rows per CTA, neighbour flags): bands of one row, ragged bands, several passes per row
    (1024, 700, 500, 0), (256, 37, 522, 0), (128, 128, 514, 0), (4096, 200, 521, 0), (12, 9, 500, 0), (64, 64, 522, 0),
    (640, 300, 514, 0), (128, 2, 500, 0), (256, 5, 541, 0), (2052, 130, 500, 0), (1024, 460, 0, 0), (512, 600, 0, 0),
])
def test_strict_steps_bit_exact_vs_oracle(gpu, pkg, orc, nx, ny, kernel, block):
    p, obst, cells0 = random_case(orc, nx, ny, seed=nx * 1000 + ny)
    iters = 7
    ref_cells, ref_av = orc.run(p, obst, iters, cells=cells0)
    with pkg.Lattice(to_param(p, iters), obst, kernel=kernel, block=block) as lat:
        lat.upload(cells0)
        assert np.array_equal(bits(lat.cells()), bits(cells0))  # upload/download round trip
        lat.run(iters)
        cells = lat.cells()
        av = lat.av_vels()
        sums, bad = lat.tot_u_sums()
        fluid = lat.fluid_cells
    assert fluid == int((obst == 0).sum())
    assert_lattice_equal(cells, ref_cells, obst)
    assert not bad.any()
    exact = exact_tot_u(orc, p, obst, cells0, iters)
    for t in range(iters):
        tot = (int(sums[t, 0]) + (int(sums[t, 1]) << 24)) * 2.0 ** -40
        assert abs(tot - exact[t][0]) <= fluid * 2.0 ** -41 + 1e-12 * exact[t][0]
        assert exact[t][1] == fluid
    np.testing.assert_allclose(av, ref_av, rtol=5e-5)


def test_division_and_sqrt_sequences_match_the_ieee_instructions(gpu, pkg):
    """2 x 2^31 quotients and 2^31 roots from the kernels' own sequences (div2_rn, speed_from_sq)
    against div.rn.f32 / sqrt.rn.f32 on the device: not one differing bit pattern."""
    for seed in (1, 2):
        assert pkg.selftest(pairs=1 << 31, seed=seed) == (0, 0)


def test_packed_collision_equals_the_scalar_cell_code(gpu, pkg):
    """The four-cell collision of the step kernels (FADD2 / FMUL2 / FFMA2, one basic block) against update_cell()
    (scalar instructions, guarded IEEE paths) on 2 x 2^24 random groups of four cells: strict flavour bit for bit
    (populations and |u|) -- in particular the toolchain has not contracted a packed multiply with a packed add."""
    for seed in (1, 2):
        assert pkg.selftest_collide("strict", sets=1 << 24, seed=seed) == (0, 0)


def test_chunked_runs_equal_one_run(gpu, pkg, orc):
    """lbm_run may be called repeatedly: 3+1+40+33 steps (graph replays and single launches, odd and
    even chunk lengths) == 77 steps, and the accelerate-at-store folding never leaks across calls."""
    p, obst, cells0 = random_case(orc, 128, 48, seed=5)
    ref_cells, ref_av = orc.run(p, obst, 77, cells=cells0)
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.upload(cells0)
        avs = []
        for n in (3, 1, 40, 33):
            lat.run(n)
            avs.append(lat.av_vels())
        assert lat.steps_done == 77
        assert_lattice_equal(lat.cells(), ref_cells, obst)
    np.testing.assert_allclose(np.concatenate(avs), ref_av, rtol=5e-5)


def test_graph_and_plain_launch_paths_agree(gpu, pkg, orc):
    p, obst, cells0 = random_case(orc, 256, 64, seed=6)
    outs = []
    for use_graph in (True, False):
        with pkg.Lattice(to_param(p), obst, use_graph=use_graph) as lat:
            lat.upload(cells0)
            lat.run(100)
            outs.append((lat.cells(), lat.tot_u_sums()[0]))
    assert np.array_equal(bits(outs[0][0]), bits(outs[1][0]))
    tot = [o[1][:, 0] + (o[1][:, 1] << 24) for o in outs]
    assert np.array_equal(tot[0], tot[1])  # integer sums: identical, not merely close


def test_av_vels_identical_across_kernel_variants(gpu, pkg, orc):
    """The |u| reduction is integer arithmetic per cell: every kernel variant / CTA shape gives the
    same sums bit for bit."""
    p, obst, cells0 = random_case(orc, 256, 40, seed=7)
    ref = None
    for kernel, block in [(0, 0), (10, 128), (21, 512), (99, 0), (99, 128), (10823, 0), (10444, 128), (11631, 0), (201, 0), (204, 0), (3000, 0), (3202, 0),
                          (3104, 0), (400, 0), (402, 0), (404, 0), (500, 0), (514, 0)]:
        with pkg.Lattice(to_param(p), obst, kernel=kernel, block=block) as lat:
            lat.upload(cells0)
            lat.run(9)
            s = lat.tot_u_sums()[0]
        s = s[:, 0] + (s[:, 1] << 24)  # the total; how it is split into the two words depends on the kernel
        ref = s if ref is None else ref
        assert np.array_equal(s, ref)


def test_state_queries_match_oracle(gpu, pkg, orc):
    p, obst, cells0 = random_case(orc, 128, 32, seed=8)
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.upload(cells0)
        lat.run(5)
        ux, uy, u, pr = lat.final_state()
        cells = lat.cells()
        av = lat.av_velocity()
        dens = lat.total_density()
        rey = lat.calc_reynolds()
    rux, ruy, ru, rpr = orc.final_state(p, cells, obst)
    for a, b in ((ux, rux), (uy, ruy), (u, ru), (pr, rpr)):
        assert np.array_equal(bits(a), bits(b))
    assert av == pytest.approx(orc.av_velocity(p, cells, obst), rel=5e-5)
    assert dens == pytest.approx(float(cells.sum(dtype=np.float64)), rel=1e-9)
    assert rey == pytest.approx(orc.calc_reynolds(p, cells, obst), rel=5e-5)


def test_total_density_conserved_without_forcing(gpu, pkg, orc):
    p, obst, cells0 = random_case(orc, 256, 64, seed=9)
    p0 = p.replace(accel=0.0)
    with pkg.Lattice(to_param(p0), obst) as lat:
        lat.upload(cells0)
        d0 = lat.total_density()
        lat.run(200)
        d1 = lat.total_density()
    assert abs(d1 / d0 - 1) < 2e-5  # fp32 rounding of 200 collisions


def test_nonfinite_cells_are_reported_not_hidden(gpu, pkg, orc):
    p, obst, cells0 = random_case(orc, 64, 16, seed=10)
    cells0[5, 7, :] = np.nan
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.upload(cells0)
        lat.run(1)
        sums, bad = lat.tot_u_sums()
        av = lat.av_vels()
    assert bad[0] >= 1 and np.isnan(av[0])


def test_invalid_arguments_are_rejected(gpu, pkg):
    from lbm_asynchronous_b200.lattice import make_param

    with pytest.raises(pkg.LbmError) as e:
        pkg.Lattice(make_param(16, 1), np.zeros((1, 16), np.int32))
    assert e.value.code == 1
    with pytest.raises(pkg.LbmError) as e:
        pkg.Lattice(make_param(16, 16), np.zeros((16, 16), np.int32), ngpus=1024)
    assert e.value.code == 2
    with pytest.raises(pkg.LbmError):
        pkg.Lattice(make_param(16, 16), np.zeros((16, 16), np.int32), halo_lag=3)
    with pkg.Lattice(make_param(16, 16), np.zeros((16, 16), np.int32)) as lat:
        with pytest.raises(pkg.LbmError):
            lat.run(-1)
        lat.run(0)
        assert lat.av_vels().size == 0


# ---------------------------------------------------------------------------------------------
# the shipped cases at their full iteration counts, vs the SerialCode binary and the goldens
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid", ["128x128", "128x256", "256x256", "1024x1024"])
def test_shipped_case_full_run_strict_is_bit_identical_to_serialcode(gpu, pkg, orc, grid):
    import hashlib

    p, obst = load_case(orc, grid)
    fx = np.load(os.path.join(GOLDEN, f"{grid}.npz"))
    with pkg.Lattice(to_param(p), obst, arith="strict") as lat:
        lat.run(p.max_iters)
        av = lat.av_vels()
        ux, uy, u, pr = lat.final_state()
        rey = lat.calc_reynolds()
    assert np.array_equal(bits(pr), bits(fx["serial_pressure"]))
    if "serial_ux" in fx:
        assert np.array_equal(bits(ux), bits(fx["serial_ux"]))
        assert np.array_equal(bits(uy), bits(fx["serial_uy"]))
        assert np.array_equal(bits(u), bits(fx["serial_u"]))
    else:
        for name, plane in (("ux", ux), ("uy", uy), ("u", u)):
            assert hashlib.sha256(plane.tobytes()).hexdigest() == str(fx[f"serial_{name}_sha256"])
    # av_vels: same per-cell values, exact sum instead of a sequential (or per-thread) fp32 sum
    np.testing.assert_allclose(av, fx["serial_av_vels"], rtol=1e-4)
    assert rey == pytest.approx(float(fx["reynolds"]), rel=1e-4)
    # check.py's rule against the reference's shipped (double precision) goldens
    a = orc.check_metric(fx["golden_av_vels"], av)
    assert np.isfinite(a) and abs(a) < 1.0
    if "golden_pressure" in fx:
        f = orc.check_metric(fx["golden_pressure"], pr.ravel())
        assert np.isfinite(f) and abs(f) < 1.0


@pytest.mark.parametrize("grid", ["128x128", "128x256", "256x256", "1024x1024"])
def test_shipped_case_full_run_fast_passes_check(gpu, pkg, orc, grid):
    p, obst = load_case(orc, grid)
    fx = np.load(os.path.join(GOLDEN, f"{grid}.npz"))
    with pkg.Lattice(to_param(p), obst, arith="fast") as lat:
        lat.run(p.max_iters)
        av = lat.av_vels()
        _, _, _, pr = lat.final_state()
    # vs the goldens: check.py's tolerance
    a = orc.check_metric(fx["golden_av_vels"], av)
    assert np.isfinite(a) and abs(a) < 1.0
    if "golden_pressure" in fx:
        assert abs(orc.check_metric(fx["golden_pressure"], pr.ravel())) < 1.0
    # vs SerialCode: fp32 re-ordering noise only (-Ofast moves av_vels by <= 0.17 %)
    assert abs(orc.check_metric(fx["serial_av_vels"], av)) < 0.5
    assert abs(orc.check_metric(fx["serial_pressure"].ravel(), pr.ravel())) < 5e-3


def test_first_steps_vs_serialcode_binary(gpu, pkg, orc):
    for grid in ("128x128", "128x256"):
        p, obst = load_case(orc, grid)
        fx = np.load(os.path.join(GOLDEN, f"steps_{grid}.npz"))
        for k in (1, 2, 3, 10, 101):
            with pkg.Lattice(to_param(p, k), obst) as lat:
                lat.run(k)
                ux, uy, u, pr = lat.final_state()
                av = lat.av_vels()
            assert np.array_equal(bits(ux), bits(fx[f"ux_{k}"]))
            assert np.array_equal(bits(uy), bits(fx[f"uy_{k}"]))
            assert np.array_equal(bits(u), bits(fx[f"u_{k}"]))
            assert np.array_equal(bits(pr), bits(fx[f"pressure_{k}"]))
            np.testing.assert_allclose(av, fx[f"av_vels_{k}"], rtol=5e-5)


# ---------------------------------------------------------------------------------------------
# the benchmark workload (BASELINE configs 4/5: synthetic channel, Bernoulli(0.005) obstacles)
# ---------------------------------------------------------------------------------------------
def test_synthetic_channel_medium_bit_exact_vs_oracle(gpu, pkg, orc):
    nx, ny, iters = 2048, 1024, 25
    obst = pkg.channel_obstacles(nx, ny)
    p = orc.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    ref_cells, ref_av = orc.run_fused(p, obst, iters)
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.run(iters)
        cells, av = lat.cells(), lat.av_vels()
    assert_lattice_equal(cells, ref_cells, obst)
    np.testing.assert_allclose(av, ref_av, rtol=1e-4)
    # the fast flavour on the same case (its own default tile shape): fp32 re-association only
    with pkg.Lattice(to_param(p), obst, arith="fast") as lat:
        lat.run(iters)
        fcells, fav = lat.cells(), lat.av_vels()
    fluid = obst == 0
    np.testing.assert_allclose(fcells[fluid], ref_cells[fluid], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(fav, ref_av, rtol=1e-3)


def test_full_size_8192_bit_exact_vs_oracle_and_slab_invariance(gpu, pkg, orc):
    """BASELINE config 4 at its full size: 20 steps of the 8192 x 8192 channel, strict lattice bit
    identical to the oracle's fused pass (OpenMP on the host cores, a few seconds), and the same
    lattice / the same integer |u| sums when the grid is cut into 3 slabs."""
    nx = ny = 8192
    iters = 20
    obst = pkg.channel_obstacles(nx, ny)
    p = orc.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    ref_cells, ref_av = orc.run_fused(p, obst, iters)
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.run(iters)
        cells, av, sums = lat.cells(), lat.av_vels(), lat.tot_u_sums()[0]
        assert lat.fluid_cells == int((obst == 0).sum())
    fluid = obst == 0
    assert np.array_equal(bits(cells[fluid]), bits(ref_cells[fluid]))
    # av_vels: the GPU's sum is exact; compare with the double-precision sum of the same per-cell values.
    # The reference's own fp32 accumulation over 67 M cells (per-thread partial sums) is only good to
    # a few 1e-3 at this size, so it is held to a loose tolerance.
    tot, n = orc.tot_u_f64(p, ref_cells, obst)
    assert av[-1] == pytest.approx(tot / n, rel=2e-6)
    np.testing.assert_allclose(av, ref_av, rtol=2e-2)
    del ref_cells
    with pkg.Lattice(to_param(p), obst, devices=[0, 0, 0]) as lat:
        lat.run(iters)
        cells3, sums3 = lat.cells(), lat.tot_u_sums()[0]
    assert np.array_equal(bits(cells3), bits(cells))
    assert np.array_equal(sums3[:, 0] + (sums3[:, 1] << 24), sums[:, 0] + (sums[:, 1] << 24))


# ---------------------------------------------------------------------------------------------
# row slabs (several slabs on one device: the halo rings, flags and lag logic without needing N GPUs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid,nslabs", [("128x128", 2), ("128x128", 5), ("128x256", 3), ("128x256", 8)])
def test_slabs_sync_equal_single_lattice(gpu, pkg, orc, grid, nslabs):
    """Sync halo mode == MPI_Waitall semantics: the decomposed run is bit-identical to the single
    lattice (and, the sums being integers, so is av_vels)."""
    p, obst = load_case(orc, grid)
    iters = 150
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.run(iters)
        one = (lat.cells(), lat.tot_u_sums()[0])
    with pkg.Lattice(to_param(p), obst, devices=[0] * nslabs) as lat:
        assert len(lat.slabs()) == nslabs
        lat.run(iters)
        many = (lat.cells(), lat.tot_u_sums()[0])
    assert np.array_equal(bits(one[0]), bits(many[0]))
    assert np.array_equal(one[1][:, 0] + (one[1][:, 1] << 24), many[1][:, 0] + (many[1][:, 1] << 24))
    ref_cells, _ = orc.run(p, obst, iters)
    assert_lattice_equal(many[0], ref_cells, obst)


def test_slabs_random_state_and_scalar_kernel(gpu, pkg, orc):
    for nx, ny, n in [(128, 23, 4), (37, 19, 3), (4, 30, 7)]:
        p, obst, cells0 = random_case(orc, nx, ny, seed=nx + ny, walls=False)  # periodic in y across the ring
        ref_cells, _ = orc.run(p, obst, 11, cells=cells0)
        with pkg.Lattice(to_param(p), obst, devices=[0] * n) as lat:
            lat.upload(cells0)
            lat.run(5)
            lat.run(6)
            assert_lattice_equal(lat.cells(), ref_cells, obst)


@pytest.mark.parametrize("lag", [2, 4])
def test_slabs_deterministic_stale_halo_equals_oracle(gpu, pkg, orc, lag):
    """halo_lag=k: boundary rows at step t use the neighbour row of step t-k (initial state while
    t < k) -- the oracle's restatement of the MPI_Testall variant's staleness model."""
    p, obst = load_case(orc, "128x128")
    iters = 90
    nslabs = 4
    starts = pkg.partition(p.ny, nslabs)
    ref_cells, ref_av = orc.run_decomposed(p, obst, starts, lag, iters)
    with pkg.Lattice(to_param(p), obst, devices=[0] * nslabs, halo_lag=lag) as lat:
        assert [s[0] for s in lat.slabs()] == starts[:-1]
        lat.run(iters)
        cells = lat.cells()
        av = lat.av_vels()
    assert_lattice_equal(cells, ref_cells, obst)
    np.testing.assert_allclose(av, ref_av, rtol=5e-5)
    exact_cells, _ = orc.run(p, obst, iters)
    assert not np.array_equal(bits(cells), bits(exact_cells))  # the lag really changes the result


def test_slab_per_process_api_on_one_rank(gpu, pkg, orc):
    """lbm_create_slab + export/connect with a ring of one (the slab is its own neighbour)."""
    p, obst, cells0 = random_case(orc, 128, 24, seed=3, walls=False)
    ref_cells, _ = orc.run(p, obst, 9, cells=cells0)
    lat = pkg.SlabLattice(to_param(p), obst, 0, p.ny, 0, 1, 0)
    with pytest.raises(pkg.LbmError):
        lat.run(1)  # not connected yet
    h = lat.export_handle()
    lat.connect(h, h)
    lat.upload(cells0)
    lat.run(9)
    assert_lattice_equal(lat.cells(), ref_cells, obst)
    lat.close()


@pytest.mark.parametrize("kernel", [401, 402, 404, 514])
def test_single_launch_kernels_give_up_when_a_neighbour_never_runs(gpu, pkg, orc, monkeypatch, kernel):
    """Two slabs of one lattice, only one of them is ever run: its resident kernel waits for the other's packets, gives up
    after the lattice's time-out (once: later waits return at once), finishes, and lbm_sync reports LBM_ETIMEOUT -- the
    device is not left spinning."""
    import time

    monkeypatch.setenv("LBM_HALO_TIMEOUT_MS", "300")
    p, obst, _ = random_case(orc, 128, 48, seed=4, walls=False)
    a = pkg.SlabLattice(to_param(p), obst[:24], 0, 24, 0, 2, 0, kernel=kernel)
    b = pkg.SlabLattice(to_param(p), obst[24:], 24, 48, 1, 2, 0, kernel=kernel)
    ha, hb = a.export_handle(), b.export_handle()
    a.connect(hb, hb)
    b.connect(ha, ha)
    t0 = time.perf_counter()
    a.run(50)
    with pytest.raises(pkg.LbmError) as ei:
        a.sync()
    assert ei.value.code == 5  # LBM_ETIMEOUT
    assert time.perf_counter() - t0 < 20.0
    a.close()
    b.close()


# ---------------------------------------------------------------------------------------------
# the host program, end to end, with check.py's rule
# ---------------------------------------------------------------------------------------------
def test_host_program_end_to_end(gpu, built, orc, tmp_path):
    exe = os.path.join(ROOT, "lbm-asynchronous_b200", "d2q9-bgk")
    grid = "128x128"
    r = subprocess.run([exe, os.path.join(INPUTS, f"input_{grid}.params"), os.path.join(INPUTS, f"obstacles_{grid}.dat")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "==done=="
    assert lines[1].startswith("Reynolds number:\t\t")
    for i, name in enumerate(["Init", "Compute", "Collate", "Total"]):
        assert lines[2 + i].startswith(f"Elapsed {name} time:\t\t\t") and lines[2 + i].endswith(" (s)")
    fx = np.load(os.path.join(GOLDEN, f"{grid}.npz"))
    assert float(lines[1].split()[-1]) == pytest.approx(float(fx["reynolds"]), rel=1e-4)
    av = orc.read_av_vels(str(tmp_path / "av_vels.dat"))
    fs = orc.read_final_state(str(tmp_path / "final_state.dat"))
    ok, a, f = orc.check_passes(fx["golden_av_vels"], av, fx["golden_pressure"], fs[:, 5])
    assert ok, (a, f)
    # the text is the reference's: same lines as SerialCode's final_state.dat would hold
    ny, nx = fx["serial_pressure"].shape
    assert fs.shape == (nx * ny, 7)
    assert np.array_equal(fs[:, 5].astype(np.float32).view(np.uint32), bits(fx["serial_pressure"]).ravel())
    assert np.array_equal(fs[:, 2].astype(np.float32).view(np.uint32), bits(fx["serial_ux"]).ravel())
    first = open(tmp_path / "final_state.dat").readline()
    assert first == "0 0 0.000000000000E+00 0.000000000000E+00 0.000000000000E+00 3.333333507180E-02 1\n"
    assert open(tmp_path / "av_vels.dat").readline().startswith("0:\t")


def test_host_program_animation_frames(gpu, built, orc, tmp_path):
    """LBM_ANIMATION_EVERY=N: the frames of write_animation_data() (SerialCode/d2q9-bgk.c:802-849), after
    timesteps tt = 0, N, 2N, ...; chunked runs leave av_vels.dat and final_state.dat unchanged."""
    exe = os.path.join(ROOT, "lbm-asynchronous_b200", "d2q9-bgk")
    grid, iters, every = "128x128", 120, 50
    tok = open(os.path.join(INPUTS, f"input_{grid}.params")).read().split()
    pf = tmp_path / "p.params"
    pf.write_text("\n".join(tok[:2] + [str(iters)] + tok[3:]) + "\n")
    of = os.path.join(INPUTS, f"obstacles_{grid}.dat")
    outs = {}
    for name, env in (("plain", {}), ("frames", {"LBM_ANIMATION_EVERY": str(every)})):
        wd = tmp_path / name
        wd.mkdir()
        r = subprocess.run([exe, str(pf), of], cwd=wd, capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        outs[name] = (open(wd / "av_vels.dat").read(), open(wd / "final_state.dat").read(), r.stdout)
    assert outs["plain"][0] == outs["frames"][0] and outs["plain"][1] == outs["frames"][1]
    p, obst = load_case(orc, grid)
    for tt in (0, 50, 100):
        assert f"Written animation data for timestep {tt}\n" in outs["frames"][2]
        lines = open(tmp_path / "frames" / "animation_data" / f"velocity_magnitude_{tt:06d}.dat").read().splitlines()
        assert lines[0] == f"# nx={p.nx} ny={p.ny} timestep={tt}"
        cells, _ = orc.run(p, obst, tt + 1)
        _, _, u, _ = orc.final_state(p, cells, obst)
        assert lines[1:] == ["%.6E" % v for v in u.ravel()]
    assert not os.path.exists(tmp_path / "frames" / "animation_data" / "velocity_magnitude_000119.dat")


def test_large_grid_with_partial_tiles_bit_exact(gpu, pkg, orc):
    """A large grid whose width is not a multiple of the 128-cell tile and whose height is not a multiple of
    the 16-row tile (partial TMA boxes in x and y, zero fill + masked stores), default kernels."""
    nx, ny, iters = 4100, 3001, 6
    obst = pkg.channel_obstacles(nx, ny, p=0.01, seed=3)
    p = orc.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    ref_cells, ref_av = orc.run_fused(p, obst, iters)
    with pkg.Lattice(to_param(p), obst) as lat:
        lat.run(iters)
        cells, av = lat.cells(), lat.av_vels()
    assert_lattice_equal(cells, ref_cells, obst)
    tot, n = orc.tot_u_f64(p, ref_cells, obst)
    assert av[-1] == pytest.approx(tot / n, rel=2e-6)


def test_packed_obstacle_map_gives_the_same_lattice(gpu, pkg, orc):
    """lbm_create_packed / lbm_create_slab_packed (one bit per cell, the device's own layout) == lbm_create from
    the reference's int map, including widths that are not a multiple of 32 and stray bits beyond nx."""
    for nx, ny in ((100, 40), (128, 33), (37, 19)):
        p, obst, cells0 = random_case(orc, nx, ny, seed=nx)
        packed = pkg.pack_obstacles(obst)
        assert packed.shape == (ny, (nx + 31) // 32) and packed.dtype == np.uint32
        if nx % 32:
            packed[:, -1] |= np.uint32(0xFFFFFFFF) << np.uint32(nx % 32)  # garbage beyond nx must be ignored
        outs = []
        for ob in (obst, packed):
            with pkg.Lattice(to_param(p), ob) as lat:
                assert lat.fluid_cells == int((obst == 0).sum())
                lat.upload(cells0)
                lat.run(9)
                outs.append(lat.cells())
        assert np.array_equal(bits(outs[0]), bits(outs[1]))
    p, obst, cells0 = random_case(orc, 128, 24, seed=3, walls=False)
    ref_cells, _ = orc.run(p, obst, 9, cells=cells0)
    lat = pkg.SlabLattice(to_param(p), pkg.pack_obstacles(obst), 0, p.ny, 0, 1, 0)
    h = lat.export_handle()
    lat.connect(h, h)
    lat.upload(cells0)
    lat.run(9)
    assert_lattice_equal(lat.cells(), ref_cells, obst)
    lat.close()
