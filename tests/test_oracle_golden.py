"""Pins the CPU oracle (oracle/pycd_oracle.c) against the reference: its shipped example
artefacts and vectors produced by running the unmodified reference (make_golden.py).
CPU only."""
import numpy as np
import pytest

import helpers as H
import oracle as O


@pytest.mark.parametrize('name,k_eff', [('hematite', 3457), ('bvo', 1557)])
def test_ewald_literal_matches_shipped_array(name, k_eff):
    ex = H.load_example(name)
    ep = H.ewald_parameters(ex)
    sc = ex.supercell
    P, keff = O.ewald_literal(H.shipped_pairwise(ex), sc.reciprocal_lattice_matrix, sc.system_volume,
                              ep.alpha, ep.r_cut, ep.k_cut, ep.dielectric, ep.k_max)
    assert keff == k_eff
    scale = np.abs(ex.P).max()
    assert np.abs(P - ex.P).max() <= 1e-13 * scale
    assert np.allclose(P, ex.P, rtol=1e-10, atol=1e-12 * scale)


@pytest.mark.parametrize('name', ['hematite', 'bvo'])
def test_min_image_matches_shipped_pair_vectors(name):
    ex = H.load_example(name)
    sc = ex.supercell
    pair = H.shipped_pairwise(ex)
    mine = O.pairwise(H.shipped_coords(ex), sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    # equal-norm images at exactly L/2 may be chosen differently: compare norms
    assert np.abs(np.linalg.norm(mine, axis=2) - np.linalg.norm(pair, axis=2)).max() < 1e-11
    same = np.abs(mine - pair).max(axis=2) < 1e-9
    assert same.mean() > 0.9
    # host-side vectorised version agrees with the C one
    host = sc.min_image(H.shipped_coords(ex)[None, :, :] - H.shipped_coords(ex)[:, None, :])
    assert np.abs(np.linalg.norm(host, axis=2) - np.linalg.norm(pair, axis=2)).max() < 1e-11


@pytest.mark.parametrize('name,n_events,literal', [('hematite', 319946, False), ('hematite', 319946, True),
                                                  ('bvo', 35201, False), ('bvo', 35201, True)])
def test_shipped_trajectory_replay_is_bit_exact(name, n_events, literal):
    ex = H.load_example(name)
    run = H.run_parameters(ex)
    rng = H.shipped_rng(ex)
    occ = run.initial_occupancy_from(rng)
    draws = H.draw_stream(rng, n_events + 64)
    res = O.KmcOracle(run, ex.P, literal=literal).trajectory(occ, draws, want_times=True)
    assert res['n_steps'] == n_events
    assert res['clamped'] == 0
    gold = H.shipped_unwrapped(ex)
    assert gold.shape == res['unwrapped'].shape
    assert np.array_equal(res['unwrapped'], gold)  # every event identical
    n, idx, val = H.shipped_time_sample(ex)
    assert n == n_events + 1
    assert np.allclose(res['times'][idx], val, rtol=1e-12, atol=0)


@pytest.mark.parametrize('tag', ['hematite_4e', 'hematite_4e_field', 'bvo_4e', 'bvo_2h'])
@pytest.mark.parametrize('literal', [False, True])
def test_reference_generated_cases(tag, literal):
    ex, z = H.load_ref_case(tag)
    run = H.run_parameters(ex)
    orc = O.KmcOracle(run, ex.P, literal=literal)
    for i in range(int(z['n_traj'])):
        rng = H.rng_from_state_bytes(z[f'rnd_state_{i}'])
        occ = run.initial_occupancy_from(rng)
        n_events = int(z[f'time_n_{i}']) - 1
        draws = H.draw_stream(rng, n_events + 16)
        res = orc.trajectory(occ, draws, want_times=True)
        if i == 0:
            assert list(occ) == list(z['occ0'])
            assert np.allclose(res['rates0'], z['rates0'], rtol=2e-12, atol=0)
            assert np.allclose(res['dg0_first'], z['dg0_0'], rtol=0, atol=1e-15)
        assert res['n_steps'] == n_events
        assert np.array_equal(res['unwrapped'], z[f'unwrapped_{i}'])
        assert np.allclose(res['times'][z[f'time_index_{i}']], z[f'time_value_{i}'], rtol=1e-12, atol=0)
        if 'drift_mobility' in z.files:
            from pycd_b200.kmc import drift_mobility
            mob = drift_mobility(res['drift'][None], run.field, run.field_mag)[0]
            assert np.allclose(mob, z['drift_mobility'][i], rtol=1e-11)


@pytest.mark.parametrize('literal', [False, True])
def test_energy_and_delg0_outputs(literal):
    """output_data energy / delg_0 (core.py:2782-2783, 2807-2809, 2826, 2855-2857)."""
    ex, z = H.load_ref_case('hematite_4e_energy')
    run = H.run_parameters(ex)
    rng = H.rng_from_state_bytes(z['rnd_state_0'])
    occ = run.initial_occupancy_from(rng)
    n_events = int(z['time_n_0']) - 1
    e0 = run.initial_energy(ex.P, occ, 0.2667)
    assert abs(e0 - z['energy_0'][0]) <= 1e-12 * abs(e0)
    res = O.KmcOracle(run, ex.P, literal=literal).trajectory(occ, H.draw_stream(rng, n_events + 8), energy0=e0)
    assert res['n_steps'] == n_events
    assert np.allclose(res['energy_grid'], z['energy_0'], rtol=1e-12, atol=0)
    assert np.allclose(res['dg0_grid'], z['delg0_0'], rtol=1e-9, atol=1e-15)
    assert np.array_equal(res['dg0_grid'] == 0, z['delg0_0'] == 0)


def test_gather_identity_equals_literal_rates():
    ex, z = H.load_ref_case('hematite_4e')
    run = H.run_parameters(ex)
    occ = list(z['occ0'])
    a = O.KmcOracle(run, ex.P, literal=False).trajectory(occ, np.full(2, 0.5))
    b = O.KmcOracle(run, ex.P, literal=True).trajectory(occ, np.full(2, 0.5))
    assert np.allclose(a['rates0'], b['rates0'], rtol=1e-12, atol=0)


@pytest.mark.parametrize('name', ['hematite', 'bvo'])
def test_msd_matches_reference_on_shipped_trajectory(name):
    from pycd_b200 import constants
    ex = H.load_example(name)
    z = np.load(H.GOLD / f'ref_msd_{name}.npz')
    sim = ex.sim
    dt = sim['time_interval'] * constants.SEC2AUTIME
    n_msd = int((sim['msd_t_final'] / constants.AUTIME2NS) / dt) + 1
    uw = H.shipped_unwrapped(ex)[None]
    res = O.msd_analysis(uw, sim['species_count'], n_msd, dt, constants.AUTIME2NS, 1 / constants.ANG2BOHR,
                         sim['trim_length'], sim['temp'], sim['n_dim'])
    assert res['msd_data'].shape == z['msd_data'].shape
    assert np.allclose(res['msd_data'], z['msd_data'], rtol=1e-11, atol=1e-9)
    want = float(str(z['log']).splitlines()[0].split('is:')[1].split()[0])
    assert f"{res['diffusivity'][0]:.3e}" == f'{want:.3e}'


def test_msd_two_trajectories_four_carriers():
    from pycd_b200 import constants
    ex, z = H.load_ref_case('hematite_4e')
    sim = ex.sim
    dt = sim['time_interval'] * constants.SEC2AUTIME
    n_msd = int((sim['msd_t_final'] / constants.AUTIME2NS) / dt) + 1
    uw = np.stack([z['unwrapped_0'], z['unwrapped_1']])
    res = O.msd_analysis(uw, sim['species_count'], n_msd, dt, constants.AUTIME2NS, 1 / constants.ANG2BOHR,
                         sim['trim_length'], sim['temp'], sim['n_dim'])
    assert np.allclose(res['msd_data'], z['msd_data'], rtol=1e-11, atol=1e-9)
    lines = str(z['msg_log']).splitlines() if 'msg_log' in z.files else str(z['msd_log']).splitlines()
    want_d = float(lines[0].split('is:')[1].split()[0])
    want_s = float(lines[1].split('is:')[1].split()[0])
    assert f"{res['diffusivity'][0]:.3e}" == f'{want_d:.3e}'
    assert f"{res['diffusivity_sem'][0]:.3e}" == f'{want_s:.3e}'


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10 (counter, key -> output)
    import ctypes as C
    u = O.philox_uniforms(0, 0, 0)
    # ctr=0,key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    a = ((0x6627e8d5 >> 5) << 26) + (0xe169c58d >> 6)
    b = ((0xbc57ac4c >> 5) << 26) + (0x9b00dbd8 >> 6)
    assert u[0] == a / 2.0 ** 53
    assert u[1] == (b + 1) / 2.0 ** 53
    assert 0.0 <= u[0] < 1.0 and 0.0 < u[1] <= 1.0


@pytest.mark.parametrize('literal', [False, True])
@pytest.mark.parametrize('tag', ['bvo_doped_init', 'bvo_doped'])
def test_doped_reference_cases(tag, literal):
    """Doping hooks of do_kmc_steps (core.py:2723-2776) against the reference's own doped runs:
    per-trajectory site_indices.npy -> shifted site energies, dopant charges, carriers started on
    dopant sites (site_charge_initiation)."""
    ex, z = H.load_ref_case(tag)
    run = H.run_parameters(ex)
    assert run.doping.active and not run.doping.pairwise_insertion
    for i in range(int(z['n_traj'])):
        dop = run.doping.load(z[f'site_indices_{i}'], run.e_rel, run.q_lat)
        assert np.array_equal(dop.e_rel, z[f'e_rel_{i}'])
        rng = H.rng_from_state_bytes(z[f'rnd_state_{i}'])
        occ = run.initial_occupancy_from(rng, dop.dopant_site_indices)
        assert list(occ) == list(z[f'occ0_{i}'])
        q = dop.q_lat(run.q_lat)
        q_all = q.copy()
        np.add.at(q_all, occ, run.q_carrier)
        assert np.array_equal(q_all, z[f'q0_{i}'])
        n_events = len(z[f'time_{i}']) - 1
        orc = O.KmcOracle(run, ex.P, literal=literal, e_rel=dop.e_rel, q_lat=q)
        res = orc.trajectory(occ, H.draw_stream(rng, n_events + 16), want_times=True)
        assert np.allclose(res['rates0'], z[f'rates0_{i}'], rtol=2e-12, atol=0)
        assert np.allclose(res['dg0_first'], z[f'dg0_0_{i}'], rtol=0, atol=1e-15)
        assert res['n_steps'] == n_events
        assert np.array_equal(res['unwrapped'], z[f'unwrapped_{i}'])
        assert np.allclose(res['times'], z[f'time_{i}'], rtol=1e-12, atol=0)
