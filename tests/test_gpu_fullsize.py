"""BASELINE config 3 at its FULL size (Hematite 10x10x10, N = 30 000, K_eff = 841 015, 64
electrons) on the GPU, checked through properties that do not need the whole reference run:

* Ewald: sampled elements against the literal oracle (core.py:799-878 evaluated for a handful of
  sites with the full k list), symmetry of the array, directly evaluated dense rows against the
  translation-expanded ones;
* KMC: the stencil step kernel against the oracle's restatement of the reference loop on the same
  Philox draws (bit-exact sites and displacement grids), stateless vs incremental rate updates,
  invisibility of sharding, conservation properties of a full-length launch.

Needs a B200 and ~16 GB of host memory for the dense array the oracle reads: `pytest -m gpu`."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from pycd_b200 import _native as nat
from pycd_b200 import ewald as EW
from pycd_b200 import kmc as K

pytestmark = pytest.mark.gpu

SIZE = [10, 10, 10]
CARRIERS = 64


@pytest.fixture(scope='module')
def cfg3(ctx):
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    ex = H.load_example('hematite')
    sim = ex.sim
    sc = Supercell(ex.lattice, SIZE, [1, 1, 1])
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), 300, 'full', 'full', sim['t_final'],
                        sim['time_interval'], [CARRIERS, 0], {}, sim['relative_energies'],
                        sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    p_unit, stats = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)
    assert sc.num_system_elements == 30000 and stats['k_eff'] == 841015
    return sc, run, ep, coords, p_unit


@pytest.fixture(scope='module')
def dense_host(ctx, cfg3):
    """The 7.2 GB array in host memory (what the oracle and the reference read)."""
    sc, run, ep, coords, p_unit = cfg3
    return EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)


def test_ewald_sampled_elements_match_the_literal_oracle(ctx, cfg3):
    """P[i, j] depends only on the pair vector, so the literal oracle evaluated on a handful of sites
    (all 1.5 M half-space k triples, no structure-factor trick) pins elements of the full array."""
    sc, run, ep, coords, p_unit = cfg3
    rng = np.random.default_rng(11)
    n = sc.num_system_elements
    sites = np.concatenate([[0, sc.n_per_cell - 1, n - 1], rng.choice(n, size=9, replace=False)])
    pair = O.pairwise(coords[sites], sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    ref, keff = O.ewald_literal(pair, sc.reciprocal_lattice_matrix, sc.system_volume, ep.alpha, ep.r_cut,
                                ep.k_cut, ep.dielectric, ep.k_max)
    assert keff == 841015
    scale = np.abs(p_unit).max()
    got = np.empty_like(ref)
    for a, s in enumerate(sites):
        row = EW.ewald_expand(ctx, sc, p_unit, int(s), int(s) + 1)[0]
        got[a] = row[sites]
    assert np.abs(got - ref).max() <= 1e-10 * scale
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * scale)


def test_ewald_symmetry_and_direct_rows(ctx, cfg3):
    """P = P^T (checked on the block of unit cell 0 and on rows taken from the far corner of the
    supercell), and rows evaluated directly (the dense, non-symmetric kernel instantiation) equal the
    translation-expanded ones."""
    sc, run, ep, coords, p_unit = cfg3
    n, m = sc.num_system_elements, sc.n_per_cell
    scale = np.abs(p_unit).max()
    assert np.abs(p_unit[:, :m] - p_unit[:, :m].T).max() <= 1e-13 * scale
    r0 = n - 2 * m - 7
    direct, stats = EW.ewald_rows(ctx, ep, coords, r0, r0 + 64)
    assert stats['rows'] == 64
    expanded = EW.ewald_expand(ctx, sc, p_unit, r0, r0 + 64)
    assert np.abs(direct - expanded).max() <= 1e-12 * scale
    # transposed elements through the expansion: P[r, c] == P[c, r]
    cols = np.array([0, 17, 12345, n - 1])
    for c in cols:
        row_c = EW.ewald_expand(ctx, sc, p_unit, int(c), int(c) + 1)[0]
        assert np.abs(row_c[r0:r0 + 64] - expanded[:, c]).max() <= 1e-13 * scale


def test_stencil_kernel_matches_oracle_at_full_size(ctx, cfg3, dense_host):
    """Same Philox draws, same initial sites: carrier sites, step counts and displacement grids of the
    lattice-stencil kernel are bit-identical to the oracle's run on the dense 30 000 x 30 000 array
    (stateless rates, and the incremental update with a refresh inside the run)."""
    sc, run, ep, coords, p_unit = cfg3
    n_traj, steps = 6, 3000
    occ = K.philox_initial_occupancy(run.tables, n_traj, CARRIERS, seed=2)
    kw = dict(dt_grid=run.time_interval / 2000, n_path=128, step_limit=steps, stop_at_grid_end=False)
    ref = O.KmcOracle(run, dense_host, rng_mode=1, seed=2, **kw).ensemble(occ)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    assert system.stencil_info()[0], system.stencil_info()
    for refresh in (1, 256):
        ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=2, refresh_interval=refresh, **kw)
        while ens.advance_resident(1024) > 0:
            pass
        got = ens.read()
        assert ens.last_kernel().startswith('kmc_step_warp_kernel<')
        ens.close()
        assert np.array_equal(got['n_steps'], ref['n_steps'])
        assert np.array_equal(got['occupancy'], ref['occupancy'])
        assert np.array_equal(got['unwrapped'], ref['unwrapped'])
        assert np.allclose(got['drift'], ref['drift'], rtol=1e-9, atol=1e-300)
    # the reference's literal O(N) dot with the charge vector, a short run of one trajectory
    lit = O.KmcOracle(run, dense_host, literal=True, rng_mode=1, seed=2,
                      **{**kw, 'step_limit': 40}).trajectory(occ[0], traj_id=0)
    gat = O.KmcOracle(run, dense_host, rng_mode=1, seed=2,
                      **{**kw, 'step_limit': 40}).trajectory(occ[0], traj_id=0)
    assert np.array_equal(lit['occupancy'], gat['occupancy'])
    assert np.allclose(lit['rates0'], gat['rates0'], rtol=1e-9, atol=0)
    system.close()


def test_full_ensemble_properties(ctx, cfg3):
    """512 trajectories x 4096 steps, the bench shape.  The event and time streams the kernel emits
    must explain its own final state: replaying the events through the neighbour table on the host
    gives the final carrier sites of all 512 trajectories, and replaying events + times through the
    reference's recording rule (core.py:2844-2861) gives the displacement grids; times increase; a
    trajectory does not depend on which ensemble it runs in."""
    sc, run, ep, coords, p_unit = cfg3
    n_traj, steps = 512, 4096
    t = run.tables
    nn = t.nn
    occ = K.philox_initial_occupancy(t, n_traj, CARRIERS, seed=2)
    n_path = 101
    kw = dict(dt_grid=run.time_interval / 500, n_path=n_path, step_limit=steps, stop_at_grid_end=False,
              rng_mode=nat.RNG_PHILOX, seed=2, refresh_interval=256)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    ens = K.KmcEnsemble(system, occ, **kw)
    res = ens.advance(steps, want_events=True, want_times=True)
    got = ens.read()
    assert ens.last_kernel().startswith('kmc_step_warp_kernel<')
    ens.close()
    assert np.all(got['n_steps'] == steps) and np.all(res['steps_done'] == steps)
    ev, tm = res['events'], res['times']
    assert ev.min() >= 0 and ev.max() < CARRIERS * nn
    assert np.all(np.diff(tm, axis=1) > 0) and np.all(tm[:, 0] > 0)
    assert np.array_equal(tm[:, -1], got['time'])

    # events -> sites, all trajectories at once
    centre = np.asarray(t.site_centre)
    neigh = np.asarray(t.neigh)
    hopvec = np.asarray(t.hopvec)
    cur = occ.copy()
    rows = np.arange(n_traj)
    n_grid = 8                                     # trajectories whose displacement grid is rebuilt
    disp = np.zeros((n_grid, CARRIERS, 3))
    row_prev = np.zeros((n_grid, CARRIERS, 3))
    grid = np.zeros((n_grid, n_path, 3 * CARRIERS))
    start = np.ones(n_grid, dtype=np.int64)        # row 0 is the zero row (core.py:2789-2791)
    for s in range(steps):
        c, slot = ev[:, s] // nn, ev[:, s] % nn
        e_old = centre[cur[rows, c]]
        assert e_old.min() >= 0
        for i in range(n_grid):
            disp[i, c[i]] += hopvec[e_old[i], slot[i]]
            end = int(tm[i, s] / kw['dt_grid'])
            if end >= start[i] + 1:
                end = min(end, n_path)
                if start[i] < n_path:
                    row_prev[i] += disp[i]
                    grid[i, start[i]:end] = row_prev[i].reshape(-1)
                    disp[i] = 0.0
                start[i] = end
        cur[rows, c] = neigh[e_old, slot]
    assert np.array_equal(cur, got['occupancy'])
    assert (start > 2).all()
    assert np.array_equal(grid, got['unwrapped'][:n_grid])
    assert np.all(centre[got['occupancy']] >= 0)

    # sharding: trajectories [100, 108) on their own
    sub = K.KmcEnsemble(system, occ[100:108], traj_id0=100, **kw)
    sub.advance_resident(steps)
    alone = sub.read()
    sub.close()
    system.close()
    assert np.array_equal(alone['occupancy'], got['occupancy'][100:108])
    assert np.array_equal(alone['unwrapped'], got['unwrapped'][100:108])
