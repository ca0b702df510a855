"""The drop-in drivers end to end on a PyCD work directory (needs a B200): same inputs, same
output files as the reference's material_setup / material_run / material_msd."""
import pickle
import shutil

import numpy as np
import pytest
import yaml

import helpers as H

pytestmark = pytest.mark.gpu


def _stage(name, tmp_path, sim_edit=None):
    src = H.GOLD / name
    dst = tmp_path / name
    shutil.copytree(src, dst)
    shutil.rmtree(dst / 'traj1')
    (dst / 'InputFiles' / 'pairwise_min_image_vector_data.npy').unlink()  # never needed here
    if sim_edit:
        cfg = yaml.safe_load(open(dst / 'simulation_parameters.yml'))
        cfg.update(sim_edit)
        yaml.safe_dump(cfg, open(dst / 'simulation_parameters.yml', 'w'))
    return dst


@pytest.mark.parametrize('name', ['hematite', 'bvo'])
def test_setup_run_msd_reproduce_the_shipped_example(tmp_path, name):
    from pycd_b200 import material_msd, material_run, material_setup
    ex = H.load_example(name)
    work = _stage(name, tmp_path)
    inp = work / 'InputFiles'
    shipped_P = np.load(inp / 'precomputed_array.npy')
    (inp / 'precomputed_array.npy').unlink()

    # --- material_setup: precomputed array + log (fresh coordinates => compare up to the
    # shipped files' site permutation of symmetry-tied O atoms, SURVEY F11: eigen-spectrum and
    # the carrier sub-block, which is order-independent for Fe / V sites)
    material_setup(inp, np.array(ex.sim['system_size']), np.array(ex.sim['pbc']), 0, 0, 1, 1, 0)
    P = np.load(inp / 'precomputed_array.npy')
    assert P.shape == shipped_P.shape
    from pycd_b200.tables import HopTables
    sites = HopTables(ex.lattice, ex.supercell, ex.hop, 'electron').sites
    sub, sub_ref = P[np.ix_(sites, sites)], shipped_P[np.ix_(sites, sites)]
    assert np.abs(sub - sub_ref).max() <= 1e-10 * np.abs(shipped_P).max()
    assert np.allclose(np.sort(np.linalg.eigvalsh((P + P.T) / 2)),
                       np.sort(np.linalg.eigvalsh((shipped_P + shipped_P.T) / 2)), rtol=0, atol=1e-9)
    log = open(inp / 'precomputed_array.log').read().splitlines()
    assert float(log[3][7:16]) == ex.cfg['alpha']
    assert any(l.startswith('Total system energy (neutral):') for l in log)

    # --- material_run on the SHIPPED array (the shipped trajectory belongs to it)
    np.save(inp / 'precomputed_array.npy', shipped_P)
    material_run(work)
    assert (work / 'Run.log').read_text().startswith('Time elapsed:')
    gold_state = pickle.load(open(ex.dir / 'traj1' / 'initial_rnd_state.dump', 'rb'))
    assert pickle.load(open(work / 'traj1' / 'initial_rnd_state.dump', 'rb')) == gold_state
    uw = np.load(work / 'traj1' / 'unwrapped_traj.npy')
    assert np.array_equal(uw, H.shipped_unwrapped(ex))
    n, idx, val = H.shipped_time_sample(ex)
    td = np.load(work / 'traj1' / 'time_data.npy')
    assert td.shape == (n,) and np.allclose(td[idx], val, rtol=1e-12, atol=0)

    # --- material_msd
    material_msd(work)
    z = np.load(H.GOLD / f'ref_msd_{name}.npz')
    msd = np.load(work / str(z['file_name']))
    assert np.allclose(msd, z['msd_data'], rtol=1e-11, atol=1e-9)
    mine = (work / str(z['file_name']).replace('MSD_Data', 'MSD_Analysis').replace('.npy', '.log')).read_text()
    assert mine.splitlines()[:2] == str(z['log']).splitlines()[:2]

    # --- idempotence: Run.log present and over_write = 0 -> nothing is rerun
    cfg = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    cfg['over_write'] = 0
    yaml.safe_dump(cfg, open(work / 'simulation_parameters.yml', 'w'))
    (work / 'traj1' / 'unwrapped_traj.npy').unlink()
    material_run(work)
    assert not (work / 'traj1' / 'unwrapped_traj.npy').exists()


def test_run_multi_trajectory_field_and_drift_mobility(tmp_path):
    """2 trajectories, 4 electrons, field on: files equal the reference run's vectors."""
    from pycd_b200 import material_run
    ex, z = H.load_ref_case('hematite_4e_field')
    work = _stage('hematite', tmp_path)
    yaml.safe_dump(yaml.safe_load(str(z['sim_yaml'])), open(work / 'simulation_parameters.yml', 'w'))
    material_run(work)
    for i in range(2):
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy'), z[f'unwrapped_{i}'])
    mob = np.loadtxt(work / 'drift_mobility.dat', ndmin=2)
    assert np.allclose(mob, z['drift_mobility'], rtol=1e-10)
    assert 'drift mobility' in (work / 'Run.log').read_text()


def test_setup_generates_loadable_neighbour_list_and_philox_run(tmp_path):
    """material_setup builds hop_neighbor_list.npy itself; a run with the optional b200 knobs
    (philox, incremental updates) produces the standard files."""
    from pycd_b200 import material_run, material_setup
    from pycd_b200.tables import load_hop_neighbor_list
    ex = H.load_example('hematite')
    work = _stage('hematite', tmp_path, {'species_count': [3, 0], 'n_traj': 3, 't_final': 2.0e-6,
                                         'b200': {'rng': 'philox', 'refresh_interval': 16,
                                                  'chunk_steps': 4096}})
    inp = work / 'InputFiles'
    (inp / 'hop_neighbor_list.npy').unlink()
    material_setup(inp, np.array([2, 2, 1]), np.array([1, 1, 1]), 1, 0, 1, 0, 0)
    hop = load_hop_neighbor_list(inp / 'hop_neighbor_list.npy')
    for hi in range(2):
        got, ref = hop['Fe:Fe'][0][hi], ex.hop['Fe:Fe'][0][hi]
        assert all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in
                   zip(got.neighbor_system_element_indices, ref.neighbor_system_element_indices))
    assert (inp / 'neighbor_list.log').exists()
    material_run(work)
    for i in range(3):
        uw = np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy')
        td = np.load(work / f'traj{i + 1}' / 'time_data.npy')
        assert uw.shape == (201, 9) and np.all(uw[0] == 0) and np.any(uw[-1] != 0)
        assert td[0] == 0.0 and np.all(np.diff(td) > 0) and td[-1] >= 2.0e-6 * 4.134137336634339e16


def test_energy_and_delg0_outputs(tmp_path):
    """output_data energy / delg_0 files equal the reference run's energy_traj.npy / delG0_traj.npy."""
    from pycd_b200 import material_run
    ex, z = H.load_ref_case('hematite_4e_energy')
    work = _stage('hematite', tmp_path)
    yaml.safe_dump(yaml.safe_load(str(z['sim_yaml'])), open(work / 'simulation_parameters.yml', 'w'))
    from pycd_b200 import ewald as EW
    from datetime import datetime
    from pycd_b200.fileio import generate_report
    generate_report(datetime.now(), work / 'InputFiles', 'precomputed_array', 1,
                    ''.join(EW.log_prefix(H.ewald_parameters(ex))))
    material_run(work)
    assert np.array_equal(np.load(work / 'traj1' / 'unwrapped_traj.npy'), z['unwrapped_0'])
    assert np.allclose(np.load(work / 'traj1' / 'energy_traj.npy'), z['energy_0'], rtol=1e-12, atol=0)
    dg = np.load(work / 'traj1' / 'delG0_traj.npy')
    assert np.allclose(dg, z['delg0_0'], rtol=1e-9, atol=1e-15)
    assert np.array_equal(dg == 0, z['delg0_0'] == 0)


def test_doped_run_consumes_preprod_site_indices(tmp_path):
    """material_run with dopants: reads each trajectory's site_indices.npy (as written by the
    reference's material_preprod) and reproduces the reference's doped trajectories."""
    from pycd_b200 import material_run
    ex, z = H.load_ref_case('bvo_doped_init')
    work = _stage('bvo', tmp_path)
    yaml.safe_dump(yaml.safe_load(str(z['sim_yaml'])), open(work / 'simulation_parameters.yml', 'w'))
    with pytest.raises(FileNotFoundError, match='site_indices.npy'):
        material_run(work)
    for i in range(2):
        (work / f'traj{i + 1}').mkdir(exist_ok=True)
        np.save(work / f'traj{i + 1}' / 'site_indices.npy', z[f'site_indices_{i}'])
    material_run(work)
    for i in range(2):
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy'), z[f'unwrapped_{i}'])
        assert np.allclose(np.load(work / f'traj{i + 1}' / 'time_data.npy'), z[f'time_{i}'], rtol=1e-12, atol=0)
