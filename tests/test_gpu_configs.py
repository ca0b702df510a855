"""BASELINE configurations 4 and 5 on the GPU (config 1 / 2 as shipped: test_gpu_parity.py,
test_gpu_drivers.py; config 2's performance variant: test_stencil_kernel_bvo_shapes_match_oracle;
config 3: test_gpu_fullsize.py).

* cfg 4 -- BVO 12x12x6 (N = 20 736, K_eff = 336 074) at FULL size: a directly evaluated dense row block
  (the sharded formulation: plan of the whole array) against the translation-expanded rows, sampled
  elements against the literal oracle (core.py:799-878, every k vector, no structure-factor trick), and
  the bench leg (bench_configs.cfg4_ewald) on a smaller cell;
* cfg 5 -- the sweep code path (pycd_b200/sweep.py: one ensemble with per-trajectory kT / field / time
  grid) at reduced size: every trajectory bit-equal to the oracle run under its own condition, and the
  per-condition diffusivities / drift mobilities equal to the analysis of the oracle's trajectories."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from pycd_b200 import _native as nat
from pycd_b200 import constants
from pycd_b200 import ewald as EW
from pycd_b200 import kmc as K
from pycd_b200 import sweep as SW

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cfg4(ctx):
    from pycd_b200.lattice import Supercell
    ex = H.load_example('bvo')
    sc = Supercell(ex.lattice, [12, 12, 6], [1, 1, 1])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    p_unit, stats = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)   # DMMA kernel over the host-built k list
    assert sc.num_system_elements == 20736 and stats['k_eff'] == 336074   # SURVEY 8, sizes table
    return sc, ep, coords, p_unit


def test_cfg4_dense_row_block_equals_translation_expanded_rows(ctx, cfg4):
    """Rows evaluated DIRECTLY with the launch plan of the whole array (what a rank of the row-sharded
    run computes: 64x128 tiles, no split-k) equal the rows expanded from unit cell 0 (32x256 tiles,
    split-k): two independent evaluation orders of the same sum."""
    sc, ep, coords, p_unit = cfg4
    n = sc.num_system_elements
    scale = np.abs(p_unit).max()
    for r0 in (0, n // 2 - 13, n - 64):
        direct, st = EW.ewald_rows(ctx, ep, coords, r0, r0 + 64, plan_rows=n)
        assert st['k_split'] == 1
        expanded = EW.ewald_expand(ctx, sc, p_unit, r0, r0 + 64)
        assert np.abs(direct - expanded).max() <= 1e-12 * scale
        assert np.allclose(direct, expanded, rtol=1e-10, atol=1e-12 * scale)
        # the block is part of a symmetric matrix
        cols = EW.ewald_expand(ctx, sc, p_unit, r0, r0 + 64)[:, r0:r0 + 64]
        assert np.abs(direct[:, r0:r0 + 64] - cols.T).max() <= 1e-12 * scale


def test_cfg4_sampled_elements_match_the_literal_oracle(ctx, cfg4):
    sc, ep, coords, p_unit = cfg4
    n = sc.num_system_elements
    rng = np.random.default_rng(4)
    sites = np.concatenate([[0, sc.n_per_cell - 1, n - 1], rng.choice(n, size=13, replace=False)])
    pair = O.pairwise(coords[sites], sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    ref, keff = O.ewald_literal(pair, sc.reciprocal_lattice_matrix, sc.system_volume, ep.alpha, ep.r_cut,
                                ep.k_cut, ep.dielectric, ep.k_max)
    assert keff == 336074
    scale = np.abs(p_unit).max()
    got = np.empty_like(ref)
    for a, s in enumerate(sites):
        direct, _ = EW.ewald_rows(ctx, ep, coords, int(s), int(s) + 1, plan_rows=n)
        got[a] = direct[0][sites]
    assert np.abs(got - ref).max() <= 1e-10 * scale
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * scale)


@pytest.mark.parametrize('example,size', [('hematite', [2, 2, 1]), ('hematite', [4, 4, 2]), ('hematite', [3, 3, 5]),
                                          ('bvo', [3, 3, 2]), ('bvo', [4, 4, 1])])
def test_class_factorised_unit_rows_equal_the_dmma_rows(ctx, example, size):
    """pycd_ewald_unit_rows (k vectors enumerated per residue class modulo the supercell size, class sums
    over the basis pairs + Fourier transform over the cells) against pycd_ewald_rows(0, n_per_cell) (DMMA
    kernel over the host-built k list): the same sum in a different order, and the same K_eff."""
    from pycd_b200.lattice import Supercell
    ex = H.load_example(example)
    sc = Supercell(ex.lattice, size, [1, 1, 1])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    rows, st_r = EW.unit_cell_rows(ctx, ep, coords, method='rows')
    cells, st_c = EW.unit_cell_rows(ctx, ep, coords, method='cells')
    assert st_c['method'] == 'cells' and st_c['k_eff'] == st_r['k_eff']
    scale = np.abs(rows).max()
    assert np.abs(cells - rows).max() <= 1e-13 * scale
    auto, st_a = EW.unit_cell_rows(ctx, ep, coords)
    assert st_a['method'] == 'cells' and np.array_equal(auto, cells)


def test_class_factorised_unit_rows_reproduce_the_shipped_arrays(ctx):
    """Through the translation expansion, the class-factorised rows give the reference's shipped
    precomputed_array.npy of both examples (the arrays are translation invariant to 3e-17, SURVEY 8 f2;
    site order of the shipped Hematite file: SURVEY F11 -- compared after sorting both to the fresh order
    is not possible, so the comparison goes through the pair distances: BVO only, whose order is stable)."""
    ex = H.load_example('bvo')
    ep = H.ewald_parameters(ex)
    sc = ex.supercell
    p_unit, st = EW.unit_cell_rows(ctx, ep, np.ascontiguousarray(sc.coordinates), method='cells')
    assert st['k_eff'] == 1557
    P = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    scale = np.abs(ex.P).max()
    # shipped BVO order = stable z sort = fresh order (F11); hop-distance degeneracy (F12) does not touch P
    assert np.abs(P - ex.P).max() <= 1e-10 * scale


def test_class_factorised_unit_rows_full_size_cfg3_and_cfg4(ctx, cfg4):
    """Full size: Hematite 10x10x10 (K_eff 841 015) and BVO 12x12x6 (336 074) -- class-factorised rows vs the
    DMMA rows, and the unsupported case (partial periodic boundaries) refused."""
    from pycd_b200.lattice import Supercell
    sc, ep, coords, p_unit = cfg4
    cells, st = EW.unit_cell_rows(ctx, ep, coords, method='cells')
    assert st['k_eff'] == 336074
    assert np.abs(cells - p_unit).max() <= 1e-13 * np.abs(p_unit).max()
    ex = H.load_example('hematite')
    sc3 = Supercell(ex.lattice, [10, 10, 10], [1, 1, 1])
    ep3 = EW.EwaldParameters(sc3, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    c3 = np.ascontiguousarray(sc3.coordinates)
    rows3, st_r = EW.unit_cell_rows(ctx, ep3, c3, method='rows')
    cells3, st_c = EW.unit_cell_rows(ctx, ep3, c3, method='cells')
    assert st_c['k_eff'] == st_r['k_eff'] == 841015
    assert np.abs(cells3 - rows3).max() <= 1e-13 * np.abs(rows3).max()
    assert st_c['fourier_ms'] < 0.25 * st_r['fourier_ms']
    scp = Supercell(ex.lattice, [3, 3, 2], [1, 1, 0])
    epp = EW.EwaldParameters(scp, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    with pytest.raises(nat.NativeError):
        EW.unit_cell_rows(ctx, epp, np.ascontiguousarray(scp.coordinates), method='cells')
    _, st_p = EW.unit_cell_rows(ctx, epp, np.ascontiguousarray(scp.coordinates))
    assert st_p['method'] == 'rows'


@pytest.mark.parametrize('nb,size', [(5, (3, 2, 4)), (30, (3, 4, 10)), (24, (2, 3, 13)), (7, (2, 2, 41)), (2, (1, 1, 3))])
def test_translation_expansion_equals_the_index_map(ctx, nb, size):
    """pycd_ewald_expand against P[i, j] = Pu[b_i, ((cell_j - cell_i) mod size) * nb + b_j] written with numpy:
    odd basis counts (8-byte path) and even ones (16-byte path), z columns shorter and longer than a CTA,
    row ranges that start and end inside a cell."""
    from types import SimpleNamespace
    sx, sy, sz = size
    n = nb * sx * sy * sz
    rng = np.random.default_rng(nb * 1000 + sz)
    pu = rng.standard_normal((nb, n))
    sc = SimpleNamespace(num_system_elements=n, system_size=size, n_per_cell=nb)
    site = np.arange(n)
    cell, b = site // nb, site % nb
    cx, cy, cz = cell // (sy * sz), (cell // sz) % sy, cell % sz
    for r0, r1 in ((0, n), (1, min(n, nb + 3)), (n - 2, n)):
        rows = np.arange(r0, r1)
        dx = (cx[None, :] - cx[rows, None]) % sx
        dy = (cy[None, :] - cy[rows, None]) % sy
        dz = (cz[None, :] - cz[rows, None]) % sz
        want = pu[b[rows, None], ((dx * sy + dy) * sz + dz) * nb + b[None, :]]
        got = EW.ewald_expand(ctx, sc, pu, r0, r1)
        assert np.array_equal(got, want)


def test_unit_cell_rows_summed_from_k_range_parts(ctx):
    """The unit-cell rows sharded by k range (bench.py / N GPUs: every rank sums one part of the k list,
    one all-reduce adds the parts): the sum of the parts equals the one-call rows to rounding, for part
    counts that do and do not divide the chunk count."""
    from pycd_b200.lattice import Supercell
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [4, 4, 2], [1, 1, 1])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    whole, st = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)
    scale = np.abs(whole).max()
    for parts in (2, 7, 8):
        acc = np.zeros_like(whole)
        for p in range(parts):
            blk, stp = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell, k_part=p, k_parts=parts)
            assert stp['k_eff'] == st['k_eff']
            acc += blk
        assert np.abs(acc - whole).max() <= 1e-13 * scale


def test_cfg4_bench_leg_on_a_small_cell(ctx):
    """The bench leg itself (rows of the rank + checks) on BVO 4x4x2."""
    import torch
    import bench_configs as BC
    res = BC.cfg4_ewald(ctx, torch.device('cuda', 0), 0, 1, None, size=(4, 4, 2), check_rows=32)
    assert res['n_sites'] == 768 and res['parity']['ok'], res['parity']
    assert res['rows_per_gpu'] == 768 and res['fp64_tflops_per_gpu'] > 0


# -------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def sweep_small(ctx):
    from pycd_b200.lattice import Supercell
    ex = H.load_example('hematite')
    sim = ex.sim
    sc = Supercell(ex.lattice, [3, 3, 2], [1, 1, 1])
    run = K.RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), 300, 'full', 'full', sim['t_final'],
                          sim['time_interval'], [24, 0], {}, sim['relative_energies'], sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    p_unit, _ = EW.ewald_rows(ctx, ep, np.ascontiguousarray(sc.coordinates), 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    return sc, run, p_unit, dense


@pytest.mark.parametrize('refresh', [1, 32])
def test_cfg5_sweep_equals_per_condition_oracle_runs(ctx, sweep_small, refresh):
    """8 conditions (T in {250, 300, 350, 400} K x field off / on) flattened into one ensemble with
    per-trajectory kT, field and time grid: every trajectory equals the oracle's run of the reference loop
    under that condition alone (same Philox key = global trajectory id), and the per-condition MSD slopes,
    diffusivities and drift mobilities equal the analysis of the oracle's trajectories."""
    sc, run, p_unit, dense = sweep_small
    conds = SW.hematite_conditions()
    n_path, n_msd, trim, per_cond, seed = 41, 21, 2, 3, 2
    intervals = SW.balanced_intervals(conds, run.n_carriers, 1200, n_path)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    res = SW.run_sweep(system, run, conds, per_cond, intervals, n_path, seed=seed, refresh_interval=refresh,
                       chunk_steps=512, n_msd=n_msd, trim=trim)
    # the same ensemble once more, to read the displacement grids themselves
    n_total = len(conds) * per_cond
    gids = np.arange(n_total)
    cond_of = gids % len(conds)
    occ = K.philox_initial_occupancy(run.tables, n_total, run.n_carriers, seed)
    ens = K.KmcEnsemble(system, occ, dt_grid=float(intervals[0]), n_path=n_path, step_limit=10 ** 12,
                        stop_at_grid_end=True, rng_mode=nat.RNG_PHILOX, seed=seed, refresh_interval=refresh,
                        kT_traj=np.array([conds[c][0] * constants.K2AUTEMP for c in cond_of]),
                        field_traj=np.array([conds[c][1] for c in cond_of]), dt_grid_traj=intervals[cond_of])
    while ens.advance_resident(512) > 0:
        pass
    got = ens.read()
    assert ens.last_kernel().startswith('kmc_step_warp_kernel')
    ens.close()
    system.close()
    assert np.array_equal(got['n_steps'].astype(float), res['n_steps'])

    uw = np.empty_like(got['unwrapped'])
    drift = np.empty_like(got['drift'])
    steps = np.empty(n_total)
    for c, (T, f) in enumerate(conds):
        orc = O.KmcOracle(run, dense, kT=T * constants.K2AUTEMP, field=f, dt_grid=float(intervals[c]), n_path=n_path,
                          stop_at_grid_end=True, rng_mode=1, seed=seed)
        for g in gids[cond_of == c]:
            one = orc.trajectory(occ[g], traj_id=int(g))
            uw[g], drift[g], steps[g] = one['unwrapped'], one['drift'], one['n_steps']
    assert np.array_equal(steps, res['n_steps'])
    assert np.array_equal(got['unwrapped'], uw), 'a sweep trajectory differs from its per-condition oracle run'
    assert np.allclose(got['drift'], drift, rtol=1e-10, atol=0)
    # hot conditions and cold ones take about the same number of steps on their own grids
    means = [r['mean_steps'] for r in res['conditions']]
    assert max(means) < 2.5 * min(means), means

    pos = uw.reshape(n_total, n_path, run.n_carriers, 3) / constants.ANG2BOHR
    avg_ref = O.msd_sd(pos, n_msd).mean(axis=2)[:, :, None]
    assert np.allclose(res['avg_sd'], avg_ref, rtol=1e-11, atol=1e-9)
    ref_rows = SW.analyse_conditions(conds, intervals, avg_ref, drift, steps, n_msd, trim)
    for a, b in zip(res['conditions'], ref_rows):
        assert a['T_K'] == b['T_K'] and a['n_traj'] == per_cond
        assert np.isclose(a['D_cm2_per_Vs'], b['D_cm2_per_Vs'], rtol=1e-9)
        assert np.isclose(a['D_sem'], b['D_sem'], rtol=1e-7, atol=1e-12 * abs(b['D_cm2_per_Vs']))
        if 'drift_mobility_cm2_per_Vs' in b:
            assert np.isclose(a['drift_mobility_cm2_per_Vs'], b['drift_mobility_cm2_per_Vs'], rtol=1e-9)
    # physics sanity of the sweep: diffusivity grows with temperature
    d_off = [r['D_cm2_per_Vs'] for r in res['conditions'] if r['field_au'][0] == 0]
    assert d_off == sorted(d_off)


def test_cfg2_bench_leg(ctx):
    """The cfg-2 bench leg (BVO 4x4x2, 16 electrons, 8 slots) at a reduced ensemble: its built-in oracle check."""
    import torch
    import bench_configs as BC
    res = BC.cfg2_bvo(ctx, torch.device('cuda', 0), 0, 1, None, traj_per_gpu=32, kmc_steps=1024, launches=2,
                      refresh=64, check_traj=6)
    assert res['kernel'] == 'kmc_step_warp_kernel<1,1,8>'
    assert res['parity']['equal'] == res['parity']['checked'] == 6, res['parity']
