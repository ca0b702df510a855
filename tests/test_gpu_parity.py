"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's
golden vectors.  Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from pycd_b200 import _native as nat
from pycd_b200 import constants
from pycd_b200 import ewald as EW
from pycd_b200 import kmc as K
from pycd_b200 import msd as M

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- Ewald ----------
@pytest.mark.parametrize('name,k_eff', [('hematite', 3457), ('bvo', 1557)])
@pytest.mark.parametrize('symmetric', [False, True])
def test_ewald_matches_shipped_array(ctx, name, k_eff, symmetric):
    """north_star tolerance: 1e-10 relative in fp64 (norm of SURVEY A.2)."""
    ex = H.load_example(name)
    ep = H.ewald_parameters(ex)
    P, stats = EW.precomputed_array(ctx, ep, coords=H.shipped_coords(ex), symmetric=symmetric)
    assert stats['k_eff'] == k_eff
    scale = np.abs(ex.P).max()
    assert np.abs(P - ex.P).max() <= 1e-10 * scale
    assert np.allclose(P, ex.P, rtol=1e-10, atol=1e-12 * scale)


def test_ewald_row_blocks_equal_full(ctx):
    ex = H.load_example('hematite')
    ep = H.ewald_parameters(ex)
    coords = H.shipped_coords(ex)
    full, _ = EW.precomputed_array(ctx, ep, coords=coords, symmetric=False)
    n = ex.supercell.num_system_elements
    for r0, r1 in ((0, 30), (30, 97), (97, n)):
        blk, _ = EW.ewald_rows(ctx, ep, coords, r0, r1)
        assert np.array_equal(blk, full[r0:r1]) or np.abs(blk - full[r0:r1]).max() < 1e-16


def test_ewald_vs_oracle_fresh_geometry_3x3x2(ctx):
    """Fresh coordinates (no shipped files): CUDA vs the literal oracle, dense and symmetric."""
    from types import SimpleNamespace
    from pycd_b200.lattice import Supercell
    ex = H.load_example('bvo')
    sc = Supercell(ex.lattice, [3, 3, 2], [1, 1, 1])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    pair = O.pairwise(sc.coordinates, sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    ref, keff = O.ewald_literal(pair, sc.reciprocal_lattice_matrix, sc.system_volume, ep.alpha, ep.r_cut,
                                ep.k_cut, ep.dielectric, ep.k_max)
    scale = np.abs(ref).max()
    for symmetric in (False, True):
        P, stats = EW.precomputed_array(ctx, ep, symmetric=symmetric)
        assert stats['k_eff'] == keff
        assert np.abs(P - ref).max() <= 1e-10 * scale
        assert np.allclose(P, ref, rtol=1e-10, atol=1e-12 * scale)


def test_ewald_partial_pbc_vs_oracle(ctx):
    from pycd_b200.lattice import Supercell
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [2, 2, 1], [1, 1, 0])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    pair = O.pairwise(sc.coordinates, sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    ref, _ = O.ewald_literal(pair, sc.reciprocal_lattice_matrix, sc.system_volume, ep.alpha, ep.r_cut,
                             ep.k_cut, ep.dielectric, ep.k_max)
    P, _ = EW.precomputed_array(ctx, ep)
    assert np.abs(P - ref).max() <= 1e-10 * np.abs(ref).max()


# ---------------------------------------------------------------- KMC ------------
def _gpu_replay(ctx, ex, run, rng, n_events_hint, chunk=32768):
    occ = run.initial_occupancy_from(rng)
    system = K.KmcSystem(ctx, run, ex.P)
    state, times, events = K.run_replay(system, [rng], np.array([occ]), chunk_steps=chunk,
                                        want_times=True, want_events=True)
    system.close()
    return occ, state, times[0], events[0]


@pytest.mark.parametrize('name,n_events', [('hematite', 319946), ('bvo', 35201)])
def test_shipped_trajectory_replay_is_bit_exact(ctx, name, n_events):
    """Given the reference's own MT19937 draws, the event sequence is identical to the
    shipped trajectory (unwrapped_traj.npy bit-equal), times to 1e-12 relative."""
    ex = H.load_example(name)
    run = H.run_parameters(ex)
    occ, state, times, events = _gpu_replay(ctx, ex, run, H.shipped_rng(ex), n_events)
    assert int(state['n_steps'][0]) == n_events
    assert int(state['clamped'][0]) == 0
    gold = H.shipped_unwrapped(ex)
    assert np.array_equal(state['unwrapped'][0], gold)
    n, idx, val = H.shipped_time_sample(ex)
    assert len(times) == n
    assert np.allclose(times[idx], val, rtol=1e-12, atol=0)
    # same events as the oracle, step by step
    rng = H.shipped_rng(ex)
    occ2 = run.initial_occupancy_from(rng)
    assert occ2 == occ
    res = O.KmcOracle(run, ex.P).trajectory(occ2, H.draw_stream(rng, n_events + 8), want_events=True)
    assert np.array_equal(res['events'], events)


@pytest.mark.parametrize('tag', ['hematite_4e', 'hematite_4e_field', 'bvo_4e', 'bvo_2h'])
def test_reference_generated_cases(ctx, tag):
    ex, z = H.load_ref_case(tag)
    run = H.run_parameters(ex)
    n_traj = int(z['n_traj'])
    rngs = [H.rng_from_state_bytes(z[f'rnd_state_{i}']) for i in range(n_traj)]
    occ = np.array([run.initial_occupancy_from(r) for r in rngs])
    assert list(occ[0]) == list(z['occ0'])
    system = K.KmcSystem(ctx, run, ex.P)
    # first-step rates against the reference's own get_process_rates
    probe = K.KmcEnsemble(system, occ[:1], rng_mode=nat.RNG_REPLAY)
    probe.advance(1, draws=np.full((1, 2), 0.5))
    rates = probe.read(unwrapped=False, rates=True)['rates'][0]
    probe.close()
    assert np.allclose(rates, z['rates0'], rtol=5e-12, atol=0)
    state, times, _ = K.run_replay(system, rngs, occ, chunk_steps=16384, want_times=True)
    for i in range(n_traj):
        assert int(state['n_steps'][i]) == int(z[f'time_n_{i}']) - 1
        assert np.array_equal(state['unwrapped'][i], z[f'unwrapped_{i}'])
        assert np.allclose(times[i][z[f'time_index_{i}']], z[f'time_value_{i}'], rtol=1e-12, atol=0)
    if 'drift_mobility' in z.files:
        mob = K.drift_mobility(state['drift'], run.field, run.field_mag)
        assert np.allclose(mob, z['drift_mobility'], rtol=1e-10)
    system.close()


@pytest.mark.parametrize('tag', ['bvo_doped_init', 'bvo_doped'])
def test_doped_reference_cases(ctx, tag):
    """Doping hooks (core.py:2723-2776) against the reference's own doped runs: per-trajectory
    site_indices.npy -> shifted site energies, dopant charges, carriers started on dopant sites;
    both trajectories of the run (different dopant sites) advance in one ensemble."""
    ex, z = H.load_ref_case(tag)
    run = H.run_parameters(ex)
    n_traj = int(z['n_traj'])
    doping = [run.doping.load(z[f'site_indices_{i}'], run.e_rel, run.q_lat) for i in range(n_traj)]
    rngs = [H.rng_from_state_bytes(z[f'rnd_state_{i}']) for i in range(n_traj)]
    occ = np.array([run.initial_occupancy_from(r, d.dopant_site_indices) for r, d in zip(rngs, doping)])
    for i in range(n_traj):
        assert list(occ[i]) == list(z[f'occ0_{i}'])
    system = K.KmcSystem(ctx, run, ex.P)
    probe = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_REPLAY, doping=doping)
    probe.advance(1, draws=np.full((n_traj, 2), 0.5))
    rates = probe.read(unwrapped=False, rates=True)['rates']
    probe.close()
    for i in range(n_traj):
        assert np.allclose(rates[i], z[f'rates0_{i}'], rtol=5e-12, atol=0)
    state, times, _ = K.run_replay(system, rngs, occ, chunk_steps=8192, want_times=True, doping=doping)
    for i in range(n_traj):
        assert int(state['n_steps'][i]) == len(z[f'time_{i}']) - 1
        assert np.array_equal(state['unwrapped'][i], z[f'unwrapped_{i}'])
        assert np.allclose(times[i], z[f'time_{i}'], rtol=1e-12, atol=0)
    system.close()


@pytest.mark.parametrize('layout', ['unit_rows', 'dense'])
@pytest.mark.parametrize('refresh', [1, 16])
def test_doped_ensemble_matches_oracle_fresh_geometry(ctx, layout, refresh):
    """Doped Philox ensemble on BVO 3x3x2 (Ewald array from the GPU): every trajectory has its own
    dopant sites / shifted shells; the checker runs one oracle per trajectory with that
    trajectory's site energies and lattice charges."""
    from pycd_b200.doping import TrajectoryDoping
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    ex = H.load_example('bvo')
    sim = ex.sim
    sc = Supercell(ex.lattice, [3, 3, 2], [1, 1, 1])
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full',
                        sim['t_final'], sim['time_interval'], [5, 0], {}, sim['relative_energies'],
                        sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    p_unit, _ = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    n_traj, steps = 6, 1500
    rng = np.random.default_rng(5)
    doping = []
    for i in range(n_traj):
        sites = rng.choice(run.tables.sites, size=1 + i % 3, replace=False)   # 1..3 dopants
        e_rel = run.e_rel.copy()
        e_rel[sites] += 0.6596 * 0.0367493
        near = np.unique(run.tables.neigh[run.tables.site_centre[sites]])
        e_rel[near] += -0.0168 * 0.0367493
        doping.append(TrajectoryDoping({'W': [int(s) for s in sites]}, e_rel, sites, np.full(len(sites), 1.0)))
    occ = K.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed=9)
    kw = dict(dt_grid=run.time_interval / 500, n_path=128, step_limit=steps, stop_at_grid_end=False)
    system = K.KmcSystem(ctx, run, p_unit if layout == 'unit_rows' else dense, layout=layout)
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=9, refresh_interval=refresh, doping=doping, **kw)
    while ens.advance_resident(496) > 0:
        pass
    got = ens.read()
    # BVO electrons in a 3x3x2 cell: 8 neighbour slots; the unit-rows layout runs the lattice-stencil kernel
    # (doped trajectories read their site energies / lattice potential per site), the dense layout the gathers
    assert ens.last_kernel() == ('kmc_step_warp_kernel<1,1,8>' if layout == 'unit_rows' else 'kmc_step_kernel')
    ens.close()
    system.close()
    for i in range(n_traj):
        ref = O.KmcOracle(run, dense, rng_mode=1, seed=9, e_rel=doping[i].e_rel, q_lat=doping[i].q_lat(run.q_lat),
                          **kw).ensemble(occ[i:i + 1], traj_id0=i)
        assert got['n_steps'][i] == ref['n_steps'][0]
        assert np.array_equal(got['occupancy'][i], ref['occupancy'][0])
        assert np.array_equal(got['unwrapped'][i], ref['unwrapped'][0])


def test_vlat_matches_oracle(ctx):
    ex = H.load_example('bvo')
    run = H.run_parameters(ex)
    system = K.KmcSystem(ctx, run, ex.P)
    assert np.allclose(system.v_lat(), O.vlat(ex.P, run.q_lat), rtol=0, atol=2e-16 * np.abs(ex.P).max() * 50)
    system.close()


def _philox_case(name='hematite', species=(8, 0), size=None):
    ex = H.load_example(name, species_count=list(species))
    return ex, H.run_parameters(ex)


@pytest.mark.parametrize('refresh', [1, 16])
def test_philox_ensemble_matches_oracle(ctx, refresh):
    """Counter-based RNG mode: 64 trajectories x 8 carriers, fixed 3000 steps; stateless and
    incremental delta-E updates give the oracle's event sequences (final sites, step counts,
    displacement grids bit-equal)."""
    ex, run = _philox_case()
    n_traj, steps = 64, 3008
    occ = K.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed=11)
    dt, n_path = run.time_interval / 200, 400
    system = K.KmcSystem(ctx, run, ex.P)
    ens = K.KmcEnsemble(system, occ, dt_grid=dt, n_path=n_path, step_limit=steps, stop_at_grid_end=False,
                        rng_mode=nat.RNG_PHILOX, seed=11, refresh_interval=refresh)
    while ens.advance_resident(1504) > 0:
        pass
    got = ens.read()
    ens.close()
    system.close()
    ref = O.KmcOracle(run, ex.P, dt_grid=dt, n_path=n_path, step_limit=steps, stop_at_grid_end=False,
                      rng_mode=1, seed=11).ensemble(occ)
    assert np.array_equal(got['n_steps'], ref['n_steps'])
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    assert np.array_equal(got['unwrapped'], ref['unwrapped'])


@pytest.mark.parametrize('refresh', [1, 16])
@pytest.mark.parametrize('case', ['hematite_3x3x2_12e', 'bvo_3x3x1_6h'])
def test_unit_rows_layout_matches_oracle_on_expanded_array(ctx, case, refresh):
    """Translation-symmetric table: the GPU gathers from the (n_per_cell, N) rows of unit
    cell 0; the oracle runs on the expanded dense (N, N) array.  Fresh geometry, Ewald
    array computed on the GPU."""
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    name, size, species = {'hematite_3x3x2_12e': ('hematite', [3, 3, 2], [12, 0]),
                           'bvo_3x3x1_6h': ('bvo', [3, 3, 1], [0, 6])}[case]
    ex = H.load_example(name)
    sim = ex.sim
    sc = Supercell(ex.lattice, size, [1, 1, 1])
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full',
                        sim['t_final'], sim['time_interval'], species, {}, sim['relative_energies'],
                        sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    p_unit, _ = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    direct, _ = EW.ewald_rows(ctx, ep, coords, 0, sc.num_system_elements)
    assert np.abs(dense - direct).max() <= 1e-13 * np.abs(direct).max()  # translation invariance
    n_traj, steps = 16, 2000
    occ = K.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed=21)
    kw = dict(dt_grid=run.time_interval / 500, n_path=128, step_limit=steps, stop_at_grid_end=False)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=21, refresh_interval=refresh, **kw)
    while ens.advance_resident(400) > 0:
        pass
    got = ens.read()
    # Hematite electrons: 4 slots; BVO holes: 12 slots, two site classes -- both on the lattice-stencil kernel
    assert system.stencil_info()[0], system.stencil_info()
    assert ens.last_kernel() == {'hematite_3x3x2_12e': 'kmc_step_warp_kernel<1,1,4>',
                                 'bvo_3x3x1_6h': 'kmc_step_warp_kernel<1,1,12>'}[case]
    # re-armed ensemble reproduces itself
    ens.reset(occ)
    while ens.advance_resident(2000) > 0:
        pass
    again = ens.read()
    ens.close()
    vl = system.v_lat()
    system.close()
    ref = O.KmcOracle(run, dense, rng_mode=1, seed=21, **kw).ensemble(occ)
    assert np.allclose(vl, O.vlat(dense, run.q_lat)[:sc.n_per_cell], rtol=0, atol=1e-14)
    assert np.array_equal(got['n_steps'], ref['n_steps'])
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    assert np.array_equal(got['unwrapped'], ref['unwrapped'])
    assert np.array_equal(again['unwrapped'], got['unwrapped'])


@pytest.mark.parametrize('refresh,chunk', [(1, 500), (8, 512)])
def test_time_grid_mode_chunked_launches(ctx, refresh, chunk):
    """Reference loop condition (run until the time grid is full) with trajectories that end
    in different launches; single carrier and multi-carrier; state carried across launches."""
    for species in ([1, 0], [5, 0]):
        ex = H.load_example('hematite', species_count=species)
        run = H.run_parameters(ex)
        n_traj = 12
        occ = K.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed=33)
        kw = dict(dt_grid=run.time_interval / 4, n_path=300, stop_at_grid_end=True)
        system = K.KmcSystem(ctx, run, ex.P)
        ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=33, refresh_interval=refresh, **kw)
        launches = 0
        while ens.advance_resident(chunk) > 0:
            launches += 1
        got = ens.read()
        ens.close()
        system.close()
        ref = O.KmcOracle(run, ex.P, rng_mode=1, seed=33, **kw).ensemble(occ)
        assert launches >= 2
        assert len(set(ref['n_steps'])) > 1
        assert np.array_equal(got['n_steps'], ref['n_steps'])
        assert np.array_equal(got['unwrapped'], ref['unwrapped'])
        assert np.all(got['time'] >= 299 * kw['dt_grid'])


def test_bad_arguments_are_rejected(ctx):
    ex = H.load_example('hematite', species_count=[2, 0])
    run = H.run_parameters(ex)
    system = K.KmcSystem(ctx, run, ex.P)
    occ = K.philox_initial_occupancy(run.tables, 2, 2, seed=1)
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, refresh_interval=16, step_limit=64,
                        stop_at_grid_end=False)
    with pytest.raises(nat.NativeError, match='multiple of refresh_interval'):
        ens.advance_resident(24)
    ens.close()
    with pytest.raises(nat.NativeError, match='never end'):
        K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, stop_at_grid_end=False, step_limit=0)
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_REPLAY)
    with pytest.raises(nat.NativeError, match='REPLAY'):
        ens.advance_resident(8)
    ens.close()
    with pytest.raises(ValueError):
        K.KmcSystem(ctx, run, ex.P[:5])
    system.close()


@pytest.fixture(scope='module')
def hematite_64e(ctx):
    """Hematite 3x3x2 (N = 540, 216 Fe), 64 electrons: the (carriers, slots) = (64, 4) shape
    served by the carrier-per-thread kernel.  Ewald array computed on the GPU."""
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    ex = H.load_example('hematite')
    sim = ex.sim
    sc = Supercell(ex.lattice, [3, 3, 2], [1, 1, 1])
    field = {'electric': {'active': 1, 'dir': [1, 0, 0], 'ld': 0, 'mag': 1e-3}}
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full',
                        sim['t_final'], sim['time_interval'], [64, 0], {}, sim['relative_energies'], field)
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    p_unit, _ = EW.ewald_rows(ctx, ep, np.ascontiguousarray(sc.coordinates), 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    return run, p_unit, dense


@pytest.mark.parametrize('layout', ['unit_rows', 'dense'])
@pytest.mark.parametrize('refresh', [1, 16])
def test_64_carriers_philox_matches_oracle(ctx, hematite_64e, layout, refresh):
    run, p_unit, dense = hematite_64e
    n_traj, steps = 24, 2400
    occ = K.philox_initial_occupancy(run.tables, n_traj, 64, seed=77)
    kw = dict(dt_grid=run.time_interval / 4000, n_path=256, step_limit=steps, stop_at_grid_end=False)
    system = K.KmcSystem(ctx, run, p_unit if layout == 'unit_rows' else dense, layout=layout)
    e0 = np.linspace(-3.0, 3.0, n_traj)
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=77, refresh_interval=refresh,
                        energy0=e0, **kw)
    while ens.advance_resident(800) > 0:
        pass
    got = ens.read(rates=True)
    eg, gg = ens.read_energy()
    if layout == 'unit_rows':   # hop vectors of the fast geometry are exactly periodic: stencil path
        assert system.stencil_info()[0], system.stencil_info()
        assert ens.last_kernel().startswith('kmc_step_warp_kernel<')
    else:
        assert ens.last_kernel().startswith('kmc_step_carrier_kernel')
    ens.close()
    system.close()
    orc = O.KmcOracle(run, dense, rng_mode=1, seed=77, **kw)
    ref = orc.ensemble(occ)
    assert np.array_equal(got['n_steps'], ref['n_steps'])
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    assert np.array_equal(got['unwrapped'], ref['unwrapped'])
    assert np.allclose(got['drift'], ref['drift'], rtol=1e-9, atol=1e-300)
    one = orc.trajectory(occ[3], traj_id=3, energy0=e0[3])
    assert np.allclose(eg[3], one['energy_grid'], rtol=1e-11, atol=1e-13)
    assert np.allclose(gg[3], one['dg0_grid'], rtol=1e-9, atol=1e-15)


def test_batch_production_api_equals_the_synchronous_calls(ctx, hematite_64e):
    """pycd_kmc_advance_async + pycd_kmc_wait (launches in flight, L2 flushes between them) and the
    double-buffered read-back (pycd_kmc_read_begin / _end with re-armed batches in between) against plain
    advance + read, and against the oracle: same sites, step counts, displacement grids; the kernel
    timers account every launch that was in flight."""
    run, p_unit, dense = hematite_64e
    n_traj, burst, n_burst = 12, 256, 3
    occ = K.philox_initial_occupancy(run.tables, n_traj, 64, seed=21)
    kw = dict(dt_grid=run.time_interval / 3000, n_path=64, step_limit=burst * n_burst, stop_at_grid_end=False)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=21, refresh_interval=16, **kw)
    for _ in range(n_burst):
        ens.advance_resident(burst)
    plain = ens.read()
    # the same three launches without a host round trip, timers settled by wait()
    ens.reset(occ, 0)
    ctx.reset_timers()
    l0 = ctx.launch_count()
    for _ in range(n_burst):
        ctx.flush_l2()
        ens.advance_async(burst)
    assert ens.wait() == 0          # step_limit reached: every trajectory finished
    assert ctx.launch_count() - l0 == n_burst
    assert ctx.class_launches(nat.KC_KMC_STEP) == n_burst and ctx.total_kernel_ms(nat.KC_KMC_STEP) > 0
    in_flight = ens.read()
    # batches A, B, A through the pipelined read-back: batch B starts from other sites
    occ_b = K.philox_initial_occupancy(run.tables, n_traj, 64, seed=22)
    bufs = [ens.pinned_buffers(), ens.pinned_buffers()]
    outs = []
    batches = (occ, occ_b, occ)

    def launch(o):
        ens.reset(o, 0)              # stream-ordered: enqueued behind whatever still runs
        for _ in range(n_burst):
            ens.advance_async(burst)
    launch(batches[0])
    for i in range(3):
        ens.read_begin(bufs[i & 1])
        if i + 1 < 3:
            launch(batches[i + 1])   # the next batch is in the queue before the host waits for this one's copy
        ens.read_end()
        outs.append({k: (None if v is None else np.array(v)) for k, v in bufs[i & 1].items()})
    assert ens.wait() == 0
    ens.close()
    system.close()
    ref = O.KmcOracle(run, dense, rng_mode=1, seed=21, **kw).ensemble(occ)
    ref_b = O.KmcOracle(run, dense, rng_mode=1, seed=21, **kw).ensemble(occ_b)
    for got, want in ((plain, ref), (in_flight, ref), (outs[0], ref), (outs[1], ref_b), (outs[2], ref)):
        assert np.array_equal(got['n_steps'], want['n_steps'])
        assert np.array_equal(got['occupancy'], want['occupancy'])
        assert np.array_equal(got['unwrapped'], want['unwrapped'])
    assert np.array_equal(plain['drift'], in_flight['drift']) and np.array_equal(plain['time'], outs[2]['time'])


@pytest.mark.parametrize('carriers', [5, 33, 64, 100])
def test_stencil_kernel_equals_gather_kernels(ctx, hematite_64e, carriers, monkeypatch):
    """The lattice-stencil kernel (one H[b_a][delta][b_y][slot] entry per carrier pair) against the
    element-gather kernels on the same unit-row table: stateless mode bit-identical (rates, times,
    events), incremental mode the same event sequence; carrier counts that leave idle threads."""
    run64, p_unit, dense = hematite_64e
    from pycd_b200.kmc import RunParameters
    ex = H.load_example('hematite')
    field = {'electric': {'active': 1, 'dir': [1, 0, 0], 'ld': 0, 'mag': 1e-3}}
    run = RunParameters(run64.lattice, run64.supercell, run64.supercell.hop_neighbor_tables(), ex.sim['temp'],
                        'full', 'full', ex.sim['t_final'], ex.sim['time_interval'], [carriers, 0], {},
                        ex.sim['relative_energies'], field)
    n_traj, steps = 8, 1600
    occ = K.philox_initial_occupancy(run.tables, n_traj, carriers, seed=5)
    kw = dict(dt_grid=run.time_interval / 4000, n_path=64, step_limit=steps, stop_at_grid_end=False)
    out = {}
    for variant in ('stencil', 'gather'):
        monkeypatch.setenv('PYCD_STENCIL', '1' if variant == 'stencil' else '0')
        system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
        assert system.stencil_info()[0] == (variant == 'stencil')
        for refresh in (1, 16):
            ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=5, refresh_interval=refresh, **kw)
            res = ens.advance(steps, want_events=True, want_times=True)
            state = ens.read(rates=True)
            assert ens.last_kernel().startswith(('kmc_step_stencil', 'kmc_step_warp')) == (variant == 'stencil')
            out[variant, refresh] = (res, state)
            ens.close()
        system.close()
    for refresh in (1, 16):
        (ra, sa), (rb, sb) = out['stencil', refresh], out['gather', refresh]
        assert np.array_equal(ra['events'], rb['events'])
        assert np.array_equal(sa['occupancy'], sb['occupancy'])
        assert np.array_equal(sa['unwrapped'], sb['unwrapped'])
        if refresh == 1:
            assert np.array_equal(sa['rates'], sb['rates'])
            assert np.allclose(ra['times'], rb['times'], rtol=1e-13, atol=0)   # k_total: tree shapes differ
            assert np.array_equal(sa['drift'], sb['drift'])
        else:
            assert np.allclose(sa['rates'], sb['rates'], rtol=1e-10, atol=0)
            assert np.allclose(ra['times'], rb['times'], rtol=1e-11, atol=0)


@pytest.mark.parametrize('variant', ['', 'stencil_1warp'])
def test_stencil_sweep_conditions_time_grid(ctx, hematite_64e, variant, monkeypatch):
    """Stencil step kernels (two warps per trajectory, and the one-warp shape large ensembles get) in
    sweep mode: per-trajectory temperature and field, the reference's loop condition (run until the
    time grid is full), state carried over several launches, incremental updates; every trajectory
    equals the checker's run of its own condition."""
    run, p_unit, dense = hematite_64e
    monkeypatch.setenv('PYCD_KMC_VARIANT', variant)
    n_traj = 6
    occ = K.philox_initial_occupancy(run.tables, n_traj, 64, seed=9)
    kTs = np.array([250, 300, 350, 400, 300, 350]) * constants.K2AUTEMP
    fields = np.zeros((n_traj, 3))
    fields[3:, 0] = 1e-3
    kw = dict(dt_grid=run.time_interval / 200, n_path=96, stop_at_grid_end=True)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=9, refresh_interval=32, kT_traj=kTs,
                        field_traj=fields, **kw)
    launches = 0
    while ens.advance_resident(256) > 0:
        launches += 1
    assert ens.last_kernel() == ('kmc_step_warp_kernel<1,2,4>' if variant else 'kmc_step_warp_kernel<2,1,4>')
    got = ens.read()
    ens.close()
    system.close()
    assert launches >= 2 and len(set(got['n_steps'])) > 1
    for i in range(n_traj):
        ref = O.KmcOracle(run, dense, kT=kTs[i], field=fields[i], rng_mode=1, seed=9, **kw).trajectory(occ[i], traj_id=i)
        assert int(got['n_steps'][i]) == ref['n_steps']
        assert np.array_equal(got['occupancy'][i], ref['occupancy'])
        assert np.array_equal(got['unwrapped'][i], ref['unwrapped'])
        assert np.allclose(got['drift'][i], ref['drift'], rtol=1e-9, atol=1e-300)


def test_64_carriers_replay_and_first_step_rates(ctx, hematite_64e):
    import random
    run, p_unit, dense = hematite_64e
    rngs = [random.Random(100 + i) for i in range(3)]
    occ = np.array([run.initial_occupancy_from(r) for r in rngs])
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    probe = K.KmcEnsemble(system, occ[:1], rng_mode=nat.RNG_REPLAY)
    probe.advance(1, draws=np.full((1, 2), 0.5))
    rates = probe.read(unwrapped=False, rates=True)['rates'][0]
    probe.close()
    orc = O.KmcOracle(run, dense, n_path=40, dt_grid=run.time_interval / 2000)
    assert np.allclose(rates, orc.trajectory(occ[0], np.full(2, 0.5))['rates0'], rtol=5e-12, atol=0)
    # replayed MT19937 streams, reference loop condition
    draws = [np.array([r.random() for _ in range(2 * 6000)]) for r in rngs]
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_REPLAY, n_path=40, dt_grid=run.time_interval / 2000)
    res = ens.advance(6000, draws=np.stack(draws), want_events=True, want_times=True)
    got = ens.read()
    ens.close()
    system.close()
    for i in range(3):
        ref = orc.trajectory(occ[i], draws[i], want_events=True, want_times=True)
        n = ref['n_steps']
        assert 0 < n < 6000 and int(res['steps_done'][i]) == n
        assert np.array_equal(res['events'][i, :n], ref['events'])
        assert np.allclose(res['times'][i, :n], ref['times'][1:], rtol=1e-12, atol=0)
        assert np.array_equal(got['unwrapped'][i], ref['unwrapped'])


def _bvo_fresh(ctx, size, species):
    """BVO supercell with fresh geometry: run parameters, unit-cell rows and the expanded dense array."""
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    ex = H.load_example('bvo')
    sim = ex.sim
    sc = Supercell(ex.lattice, size, [1, 1, 1])
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full',
                        sim['t_final'], sim['time_interval'], species, {}, sim['relative_energies'],
                        sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    p_unit, _ = EW.ewald_rows(ctx, ep, np.ascontiguousarray(sc.coordinates), 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    return run, p_unit, dense


@pytest.mark.parametrize('refresh', [1, 16])
@pytest.mark.parametrize('case', ['cfg2_perf_4x4x2_16e', 'bvo_3x3x2_40e', 'bvo_3x3x2_40h'])
def test_stencil_kernel_bvo_shapes_match_oracle(ctx, case, refresh):
    """BASELINE config 2, performance variant (BVO 4x4x2, N = 768, 8 neighbours per V, 16
    electrons, fresh geometry) and the two-warp shapes of the 8-slot (V:V electrons) and 12-slot (O:O holes,
    two site classes, core.py:1961-1970) stencil kernels against the oracle on the expanded dense array."""
    size, species, kernel = {'cfg2_perf_4x4x2_16e': ([4, 4, 2], [16, 0], 'kmc_step_warp_kernel<1,1,8>'),
                             'bvo_3x3x2_40e': ([3, 3, 2], [40, 0], 'kmc_step_warp_kernel<2,1,8>'),
                             'bvo_3x3x2_40h': ([3, 3, 2], [0, 40], 'kmc_step_warp_kernel<2,1,12>')}[case]
    run, p_unit, dense = _bvo_fresh(ctx, size, species)
    assert run.tables.nn == (8 if species[0] else 12)
    n_traj, steps = 12, 1600
    occ = K.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed=31)
    kw = dict(dt_grid=run.time_interval / 2000, n_path=96, step_limit=steps, stop_at_grid_end=False)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    assert system.stencil_info()[0], system.stencil_info()
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=31, refresh_interval=refresh, **kw)
    res = ens.advance(steps, want_events=True)
    got = ens.read()
    assert ens.last_kernel() == kernel
    ens.close()
    system.close()
    orc = O.KmcOracle(run, dense, rng_mode=1, seed=31, **kw)
    ref = orc.ensemble(occ)
    assert np.array_equal(got['n_steps'], ref['n_steps'])
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    assert np.array_equal(got['unwrapped'], ref['unwrapped'])
    one = orc.trajectory(occ[5], traj_id=5, cap_steps=steps, want_events=True)
    assert np.array_equal(res['events'][5, :steps], one['events'][:steps])


@pytest.mark.parametrize('refresh', [1, 32])
def test_stencil_kernel_100_carriers_matches_oracle(ctx, hematite_64e, refresh):
    """More than 64 carriers: two warps, two carriers per lane (kmc_step_warp_kernel<2,2,4>), field on."""
    run64, p_unit, dense = hematite_64e
    from pycd_b200.kmc import RunParameters
    ex = H.load_example('hematite')
    field = {'electric': {'active': 1, 'dir': [1, 0, 0], 'ld': 0, 'mag': 1e-3}}
    run = RunParameters(run64.lattice, run64.supercell, run64.supercell.hop_neighbor_tables(), ex.sim['temp'],
                        'full', 'full', ex.sim['t_final'], ex.sim['time_interval'], [100, 0], {},
                        ex.sim['relative_energies'], field)
    n_traj, steps = 6, 1280
    occ = K.philox_initial_occupancy(run.tables, n_traj, 100, seed=13)
    kw = dict(dt_grid=run.time_interval / 8000, n_path=64, step_limit=steps, stop_at_grid_end=False)
    for with_field in (True, False):   # featured and plain variants
        system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows', field=None if with_field else np.zeros(3))
        ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=13, refresh_interval=refresh, **kw)
        while ens.advance_resident(640) > 0:
            pass
        got = ens.read()
        assert ens.last_kernel() == 'kmc_step_warp_kernel<2,2,4>'
        ens.close()
        system.close()
        ref = O.KmcOracle(run, dense, field=None if with_field else np.zeros(3), rng_mode=1, seed=13, **kw).ensemble(occ)
        assert np.array_equal(got['n_steps'], ref['n_steps'])
        assert np.array_equal(got['occupancy'], ref['occupancy'])
        assert np.array_equal(got['unwrapped'], ref['unwrapped'])
        if with_field:
            assert np.allclose(got['drift'], ref['drift'], rtol=1e-9, atol=1e-300)


def test_sharding_is_invisible(ctx):
    """Trajectories keyed by global id: running [0,32) and [32,64) separately equals [0,64)."""
    ex, run = _philox_case(species=(4, 0))
    occ = K.philox_initial_occupancy(run.tables, 64, run.n_carriers, seed=3)
    system = K.KmcSystem(ctx, run, ex.P)

    def go(lo, hi):
        ens = K.KmcEnsemble(system, occ[lo:hi], step_limit=2000, stop_at_grid_end=False, n_path=64,
                            dt_grid=run.time_interval / 50, rng_mode=nat.RNG_PHILOX, seed=3, traj_id0=lo)
        while ens.advance_resident(1000) > 0:
            pass
        out = ens.read()
        ens.close()
        return out
    whole, a, b = go(0, 64), go(0, 32), go(32, 64)
    assert np.array_equal(whole['unwrapped'], np.concatenate([a['unwrapped'], b['unwrapped']]))
    assert np.array_equal(whole['occupancy'], np.concatenate([a['occupancy'], b['occupancy']]))
    system.close()


def test_per_trajectory_conditions(ctx):
    """Sweep mode: per-trajectory kT / field equal separate single-condition runs."""
    ex, run = _philox_case(species=(4, 0))
    occ = K.philox_initial_occupancy(run.tables, 8, run.n_carriers, seed=5)
    kTs = np.array([250, 300, 350, 400] * 2) * constants.K2AUTEMP
    fields = np.zeros((8, 3))
    fields[4:, 0] = 1e-4
    system = K.KmcSystem(ctx, run, ex.P)
    kw = dict(step_limit=1500, stop_at_grid_end=False, n_path=32, dt_grid=run.time_interval / 20,
              rng_mode=nat.RNG_PHILOX, seed=5)
    ens = K.KmcEnsemble(system, occ, kT_traj=kTs, field_traj=fields, **kw)
    ens.advance_resident(1500)
    got = ens.read()
    ens.close()
    system.close()
    for i in range(8):
        ref = O.KmcOracle(run, ex.P, kT=kTs[i], field=fields[i], step_limit=1500, stop_at_grid_end=False,
                          n_path=32, dt_grid=run.time_interval / 20, rng_mode=1, seed=5
                          ).trajectory(occ[i], traj_id=i)
        assert np.array_equal(got['occupancy'][i], ref['occupancy'])
        assert np.array_equal(got['unwrapped'][i], ref['unwrapped'])
        assert np.allclose(got['drift'][i], ref['drift'], rtol=1e-10, atol=1e-300)


def test_many_processes_per_thread(ctx):
    """n_proc > 256 (several processes per thread): 80 electrons on Hematite 2x2x1... the
    sublattice has 48 sites, carriers may share sites (no exclusion)."""
    ex = H.load_example('hematite', species_count=[40, 0])
    run = H.run_parameters(ex)
    run.n_carriers = 80
    run.n_proc = 80 * run.tables.nn
    rng = np.random.default_rng(0)
    occ = run.tables.sites[rng.integers(0, run.tables.n_centres, size=(4, 80))].astype(np.int32)
    system = K.KmcSystem(ctx, run, ex.P)
    kw = dict(step_limit=400, stop_at_grid_end=False, n_path=16, dt_grid=run.time_interval / 100)
    ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=1, **kw)
    ens.advance_resident(400)
    got = ens.read()
    ens.close()
    system.close()
    ref = O.KmcOracle(run, ex.P, rng_mode=1, seed=1, **kw).ensemble(occ)
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    assert np.array_equal(got['unwrapped'], ref['unwrapped'])


# ---------------------------------------------------------------- MSD ------------
@pytest.mark.parametrize('name', ['hematite', 'bvo'])
def test_msd_matches_reference_on_shipped_trajectory(ctx, name):
    ex = H.load_example(name)
    z = np.load(H.GOLD / f'ref_msd_{name}.npz')
    sim = ex.sim
    mp = M.MsdParameters(sim['n_dim'], sim['species_count'], 1, sim['t_final'], sim['time_interval'],
                         sim['msd_t_final'], sim['trim_length'], sim['temp'], sim['repr_time'], sim['repr_dist'])
    uw = H.shipped_unwrapped(ex)[None]
    avg = M.species_avg_sd(ctx, uw, 1, mp.n_path, mp.total_species, mp.n_msd, mp.dist_conversion, mp.type_offsets)
    res = M.analyse(mp, avg)
    assert np.allclose(res['msd_data'], z['msd_data'], rtol=1e-11, atol=1e-9)
    want = float(str(z['log']).splitlines()[0].split('is:')[1].split()[0])
    assert f"{res['diffusivity'][0]:.3e}" == f'{want:.3e}'


def test_msd_multi_trajectory_multi_carrier(ctx):
    ex, z = H.load_ref_case('hematite_4e')
    sim = ex.sim
    mp = M.MsdParameters(sim['n_dim'], sim['species_count'], 2, sim['t_final'], sim['time_interval'],
                         sim['msd_t_final'], sim['trim_length'], sim['temp'], sim['repr_time'], sim['repr_dist'])
    uw = np.stack([z['unwrapped_0'], z['unwrapped_1']])
    avg, car = M.species_avg_sd(ctx, uw, 2, mp.n_path, mp.total_species, mp.n_msd, mp.dist_conversion,
                                mp.type_offsets, want_carrier=True)
    pos = uw.reshape(2, mp.n_path, 4, 3) * mp.dist_conversion
    assert np.allclose(car, O.msd_sd(pos, mp.n_msd), rtol=1e-12, atol=1e-12)
    res = M.analyse(mp, avg)
    assert np.allclose(res['msd_data'], z['msd_data'], rtol=1e-11, atol=1e-9)
    lines = str(z['msd_log']).splitlines()
    assert f"{res['diffusivity'][0]:.3e}" == f"{float(lines[0].split('is:')[1].split()[0]):.3e}"
    assert f"{res['diffusivity_sem'][0]:.3e}" == f"{float(lines[1].split('is:')[1].split()[0]):.3e}"


def test_errors_are_loud(ctx):
    ex = H.load_example('hematite')
    run = H.run_parameters(ex)
    system = K.KmcSystem(ctx, run, ex.P)
    with pytest.raises(nat.NativeError):
        K.KmcEnsemble(system, np.array([[10 ** 6]]))  # out of range
    with pytest.raises(nat.NativeError):
        K.KmcEnsemble(system, np.array([[ex.supercell.num_system_elements - 1]]))  # an O site
    system.close()
