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


def _fast_path_workdir(tmp_path, n_traj=16, size=(4, 4, 4), carriers=64, extra_b200=None):
    """A fresh PyCD work directory for the ensemble shape of BASELINE config 3 at reduced size: Hematite
    4x4x4 (N = 1920), 64 electrons, Philox streams."""
    work = _stage('hematite', tmp_path, {
        'system_size': list(size), 'species_count': [carriers, 0], 'n_traj': n_traj, 't_final': 1.0e-7,
        'time_interval': 1.0e-9, 'msd_t_final': 50, 'trim_length': 5, 'random_seed': 7,
        'b200': dict({'rng': 'philox', 'chunk_steps': 4096}, **(extra_b200 or {}))})
    inp = work / 'InputFiles'
    for f in ('hop_neighbor_list.npy', 'precomputed_array.npy', 'precomputed_array.log'):
        (inp / f).unlink(missing_ok=True)
    return work, inp


def _oracle_of_workdir(work, inp):
    """The checker's run of the same work directory: oracle on the dense array material_setup wrote."""
    import oracle as O
    from pycd_b200 import kmc as K
    from pycd_b200.config import load_material_parameters, load_simulation_parameters
    from pycd_b200.lattice import Lattice, Supercell
    from pycd_b200.tables import load_hop_neighbor_list
    sim = load_simulation_parameters(work)
    lat = Lattice(load_material_parameters(inp))
    sc = Supercell(lat, sim['system_size'], sim['pbc'])
    run = K.RunParameters(lat, sc, load_hop_neighbor_list(inp / 'hop_neighbor_list.npy'), sim['temp'],
                          sim['ion_charge_type'], sim['species_charge_type'], sim['t_final'], sim['time_interval'],
                          sim['species_count'], sim['initial_occupancy'], sim['relative_energies'],
                          sim['external_field'], sim.get('doping'))
    occ = K.philox_initial_occupancy(run.tables, int(sim['n_traj']), run.n_carriers, int(sim['random_seed']))
    dense = np.load(inp / 'precomputed_array.npy')
    ref = O.KmcOracle(run, dense, rng_mode=1, seed=int(sim['random_seed']), stop_at_grid_end=True).ensemble(occ)
    return run, ref


def test_drivers_reach_the_stencil_kernel_and_match_the_oracle(tmp_path):
    """setup -> run -> msd through PyCD's entry points on a full-PBC supercell: material_setup writes the
    unit-cell rows, material_run picks them up by itself, runs the lattice-stencil kernel with incremental
    updates (default refresh interval) and writes trajectories equal to the oracle's; material_msd equals
    the oracle's analysis of the same trajectories."""
    import json
    import oracle as O
    from pycd_b200 import material_msd, material_run, material_setup
    from pycd_b200.material_setup import UNIT_ROWS_FILE
    work, inp = _fast_path_workdir(tmp_path)
    material_setup(inp, np.array([4, 4, 4]), np.array([1, 1, 1]), 1, 0, 1, 0, 0)
    assert np.load(inp / UNIT_ROWS_FILE).shape == (30, 1920)
    material_run(work)
    info = json.load(open(work / 'Run.b200.json'))
    assert info['p_layout'] == 'unit_rows' and info['stencil'] and info['refresh_interval'] == 256
    assert info['step_kernel'] == 'kmc_step_warp_kernel<2,1,4>'
    run, ref = _oracle_of_workdir(work, inp)
    assert info['kmc_steps'] == int(ref['n_steps'].sum())
    for i in range(16):
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy'), ref['unwrapped'][i])
    material_msd(work)
    sim = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    from pycd_b200 import constants
    from pycd_b200.msd import MsdParameters
    # step counts exactly as Analysis.__init__ truncates them (core.py:2968-2972: 50 lags here, not 51)
    mp = MsdParameters(sim['n_dim'], [64, 0], 16, sim['t_final'], sim['time_interval'], sim['msd_t_final'],
                       sim['trim_length'], sim['temp'])
    chk = O.msd_analysis(ref['unwrapped'], [64, 0], mp.n_msd, run.time_interval, constants.AUTIME2NS,
                         1 / constants.ANG2BOHR, 5, sim['temp'], sim['n_dim'])
    msd = np.load(next(work.glob('MSD_Data_*.npy')))
    assert np.allclose(msd, chk['msd_data'], rtol=1e-11, atol=1e-9)
    log = next(work.glob('MSD_Analysis_*.log')).read_text()
    assert f"{chk['diffusivity'][0]:.3e}" in log


def _run_sharded(work, n_ranks, backend, shared_gpu):
    import os
    import subprocess
    import sys
    env = dict(os.environ, PYCD_DIST_BACKEND=backend)
    if shared_gpu:
        env['PYCD_B200_DEVICE'] = '0'
    port = 29600 + os.getpid() % 300
    code = ("import sys; sys.path.insert(0, %r); from pathlib import Path; from pycd_b200 import material_run; "
            "material_run(Path(%r))" % (str(H.GOLD.parents[1]), str(work)))
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n_ranks}',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), '--no-python',
                          sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.parametrize('mode', ['gloo_shared_gpu', 'nccl_two_gpus'])
def test_material_run_under_torchrun_equals_single_process(tmp_path, mode):
    """material_run launched as two ranks (trajectory blocks [0, 5) and [5, 10)): every traj<i>/ file, the
    drift mobility and the step counts equal those of the one-process run of the same directory.  On a
    one-GPU box both ranks share the GPU and gather over gloo; with two GPUs they use NCCL."""
    import json
    import torch
    from pycd_b200 import material_run, material_setup
    if mode == 'nccl_two_gpus' and torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    field = {'electric': {'active': 1, 'dir': [1, 0, 0], 'ld': 0, 'mag': 1.0e-4}}
    work, inp = _fast_path_workdir(tmp_path, n_traj=10, size=(3, 3, 2), carriers=40, extra_b200={'refresh_interval': 64})
    cfg = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    cfg['external_field'] = field
    yaml.safe_dump(cfg, open(work / 'simulation_parameters.yml', 'w'))
    material_setup(inp, np.array([3, 3, 2]), np.array([1, 1, 1]), 1, 0, 1, 0, 0)
    material_run(work)
    single = [np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy') for i in range(10)]
    single_t = [np.load(work / f'traj{i + 1}' / 'time_data.npy') for i in range(10)]
    single_mob = np.loadtxt(work / 'drift_mobility.dat', ndmin=2)
    steps = json.load(open(work / 'Run.b200.json'))['kmc_steps']
    for i in range(10):
        shutil.rmtree(work / f'traj{i + 1}')
    (work / 'Run.log').unlink()
    (work / 'drift_mobility.dat').unlink()
    _run_sharded(work, 2, 'gloo' if mode == 'gloo_shared_gpu' else 'nccl', mode == 'gloo_shared_gpu')
    infos = [json.load(open(work / f'Run.b200.rank{r}.json')) for r in range(2)]
    assert [i['trajectories'] for i in infos] == [[0, 5], [5, 10]]
    assert sum(i['kmc_steps'] for i in infos) == steps
    assert all(i['step_kernel'].startswith('kmc_step_warp_kernel') for i in infos)
    if mode == 'nccl_two_gpus':
        assert sorted(i['device'] for i in infos) == [0, 1]
    for i in range(10):
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy'), single[i])
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'time_data.npy'), single_t[i])
    assert np.array_equal(np.loadtxt(work / 'drift_mobility.dat', ndmin=2), single_mob)
    assert 'drift mobility' in (work / 'Run.log').read_text()


def test_philox_run_honours_explicit_initial_sites(tmp_path):
    """b200.rng: philox draws only the carriers the YAML did not place (initial_occupancy, core.py:2498-2510)."""
    from pycd_b200 import material_run
    from pycd_b200.tables import HopTables
    ex = H.load_example('hematite')
    sites = HopTables(ex.lattice, ex.supercell, ex.hop, 'electron').sites
    pinned = [int(sites[3]), int(sites[17])]
    work = _stage('hematite', tmp_path, {'species_count': [3, 0], 'n_traj': 2, 't_final': 2.0e-9,
                                         'time_interval': 1.0e-9, 'initial_occupancy': {'electron': pinned},
                                         'b200': {'rng': 'philox', 'chunk_steps': 64, 'refresh_interval': 1}})
    material_run(work)
    # with ~3 carriers x 3.2e9 hops/s, 2 ns hold ~20 hops: replay them backwards from the final positions is not
    # needed -- the displacement of a carrier is a sum of hop vectors from its START site, so check the start
    # sites through the oracle run of the same occupancy instead
    import oracle as O
    from pycd_b200 import kmc as K
    from pycd_b200.material_run import PhiloxSampler
    run = H.run_parameters(ex, dict(ex.sim, species_count=[3, 0], t_final=2.0e-9, time_interval=1.0e-9,
                                    initial_occupancy={'electron': pinned}))
    occ = np.array([run.initial_occupancy_from(PhiloxSampler(ex.sim['random_seed'], i)) for i in range(2)])
    assert all(list(o[:2]) == pinned and o[2] not in pinned for o in occ)
    ref = O.KmcOracle(run, ex.P, rng_mode=1, seed=ex.sim['random_seed'], stop_at_grid_end=True).ensemble(occ)
    for i in range(2):
        assert np.array_equal(np.load(work / f'traj{i + 1}' / 'unwrapped_traj.npy'), ref['unwrapped'][i])
