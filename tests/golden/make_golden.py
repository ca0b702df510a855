"""Generates tests/golden/ from the reference (run in the build container only, where
/root/reference exists):

  * copies the reference's shipped example artefacts (inputs + golden outputs; data, not
    source) that pin the hot path, with time_data decimated;
  * runs the UNMODIFIED reference (through oracle/ref_harness.py's compatibility shims)
    on extra cases -- multi-carrier, field on, BVO holes (two site classes), 2-trajectory
    MSD -- and stores its outputs as vectors.

    python tests/golden/make_golden.py
"""
import pickle
import random as rnd
import shutil
import sys
import time
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / 'oracle'))
import ref_harness as rh  # noqa: E402

GOLD = Path(__file__).resolve().parent
REF = rh.REFERENCE_ROOT
TMP = Path('/tmp/pycd_golden_work')


def sample_indices(n):
    idx = np.unique(np.concatenate([np.arange(min(n, 2048)), np.arange(0, n, 64),
                                    np.arange(max(0, n - 2048), n)]))
    return idx


def copy_example(name):
    src = REF / 'examples' / name
    dst = GOLD / name.lower()
    if dst.exists():
        shutil.rmtree(dst)
    (dst / 'InputFiles').mkdir(parents=True)
    (dst / 'traj1').mkdir()
    for f in ('POSCAR', 'sys_config.yml', 'hop_neighbor_list.npy', 'precomputed_array.npy',
              'pairwise_min_image_vector_data.npy'):
        shutil.copy(src / 'InputFiles' / f, dst / 'InputFiles' / f)
    cfg = yaml.safe_load(open(src / 'simulation_parameters.yml'))
    cfg['output_data']['hdf5_output']['write'] = 0
    cfg['output_data']['hdf5_output']['enabled'] = False
    yaml.safe_dump(cfg, open(dst / 'simulation_parameters.yml', 'w'))
    shutil.copy(src / 'traj1' / 'initial_rnd_state.dump', dst / 'traj1' / 'initial_rnd_state.dump')
    uw = np.load(src / 'traj1' / 'unwrapped_traj.npy')
    np.savez_compressed(dst / 'traj1' / 'unwrapped_traj.npz', unwrapped=uw)
    td = np.load(src / 'traj1' / 'time_data.npy')
    idx = sample_indices(len(td))
    np.savez_compressed(dst / 'traj1' / 'time_data_sample.npz', n=len(td), index=idx, value=td[idx])
    for p in dst.rglob('*'):
        p.chmod(0o755 if p.is_dir() else 0o644)
    return dst


def reference_run_object(work):
    """Builds the reference's Run exactly like PyCD/material_run.py:13-150 does."""
    core = rh.load_reference()
    sim = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    inp = work / 'InputFiles'
    params = yaml.safe_load(open(inp / 'sys_config.yml'))
    params['input_coord_file_location'] = inp / 'POSCAR'
    mat = core.Material(core.ReturnValues(**params))
    nb = core.Neighbors(mat, np.asarray(sim['system_size']), np.asarray(sim['pbc']))
    hop = np.load(inp / 'hop_neighbor_list.npy')[()]
    pair = np.load(inp / 'pairwise_min_image_vector_data.npy')
    alpha = None
    for i, line in enumerate(open(inp / 'precomputed_array.log')):
        if i == 3:
            alpha = float(line[7:16])
    system = core.System(mat, nb, hop, pair, alpha, params['r_cut'], params['k_cut'],
                         params['precision_parameters'], [], [])
    P = np.load(inp / 'precomputed_array.npy')
    run = core.Run(system, P, sim['temp'], sim['ion_charge_type'], sim['species_charge_type'],
                   sim['n_traj'], sim['t_final'], sim['time_interval'],
                   np.asarray(sim['species_count']), sim['initial_occupancy'],
                   sim['relative_energies'], sim['external_field'], sim['doping'])
    return core, run, sim


def reference_case(tag, example, species_count, extra, with_msd=False):
    """material_run (+ material_msd) of the unmodified reference -> vectors."""
    core = rh.load_reference()
    from PyCD.material_run import material_run
    work = rh.stage_example(example, TMP / tag, species_count=species_count, extra=extra)
    t0 = time.time()
    material_run(work)
    dt = time.time() - t0
    sim = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    n_traj = int(sim['n_traj'])
    out = {'species_count': np.asarray(species_count), 'n_traj': n_traj,
           'sim_yaml': yaml.safe_dump(sim), 'example': example, 'ref_seconds': dt}
    total_steps = 0
    for i in range(n_traj):
        d = work / f'traj{i + 1}'
        out[f'rnd_state_{i}'] = np.frombuffer((d / 'initial_rnd_state.dump').read_bytes(), dtype=np.uint8)
        out[f'unwrapped_{i}'] = np.load(d / 'unwrapped_traj.npy')
        td = np.load(d / 'time_data.npy')
        idx = sample_indices(len(td))
        out[f'time_n_{i}'] = len(td)
        out[f'time_index_{i}'] = idx
        out[f'time_value_{i}'] = td[idx]
        total_steps += len(td) - 1
        for key, fname in (('energy', 'energy_traj.npy'), ('delg0', 'delG0_traj.npy')):
            if (d / fname).exists():
                out[f'{key}_{i}'] = np.load(d / fname)
    if (work / 'drift_mobility.dat').exists():
        out['drift_mobility'] = np.loadtxt(work / 'drift_mobility.dat', ndmin=2)
    # first-step rates from the reference's own rate routine
    core, run, _ = reference_run_object(work)
    rnd_mod = core.rnd
    rnd_mod.setstate(pickle.load(open(work / 'traj1' / 'initial_rnd_state.dump', 'rb')))
    occ = run.generate_initial_occupancy({})
    q = run.charge_config(occ, {})
    attrs = run.get_process_attributes(occ)
    k_list, dg0, hopv = run.get_process_rates(attrs, q)
    out['occ0'] = np.asarray(occ)
    out['rates0'] = np.asarray(k_list)
    out['dg0_0'] = np.asarray(dg0)
    out['hopvec0'] = np.asarray(hopv)
    out['new_sites0'] = np.asarray(attrs[1])
    if with_msd:
        from PyCD.material_msd import material_msd
        material_msd(work)
        f = sorted(work.glob('MSD_Data_*.npy'))[0]
        out['msd_data'] = np.load(f)
        out['msd_file_name'] = f.name
        out['msd_log'] = sorted(work.glob('MSD_Analysis_*.log'))[0].read_text()
    print(f'{tag}: {total_steps} steps in {dt:.1f} s ({total_steps / dt:.0f} steps/s, reference CPU)')
    np.savez_compressed(GOLD / f'ref_{tag}.npz', **out)


def shipped_msd(example):
    """Analysis.compute_msd of the reference on the shipped trajectory."""
    rh.load_reference()
    from PyCD.material_msd import material_msd
    work = rh.stage_example(example, TMP / f'msd_{example}', with_log=False)
    (work / 'traj1').mkdir(exist_ok=True)
    shutil.copy(REF / 'examples' / example / 'traj1' / 'unwrapped_traj.npy', work / 'traj1')
    material_msd(work)
    f = sorted(work.glob('MSD_Data_*.npy'))[0]
    np.savez_compressed(GOLD / f'ref_msd_{example.lower()}.npz', msd_data=np.load(f), file_name=f.name,
                        log=sorted(work.glob('MSD_Analysis_*.log'))[0].read_text())
    print(f'msd {example}:', sorted(work.glob('MSD_Analysis_*.log'))[0].read_text().splitlines()[:2])


def energy_case():
    """energy_traj.npy / delG0_traj.npy of the reference (output_data energy / delg_0 on)."""
    reference_case('hematite_4e_energy', 'Hematite', [4, 0],
                   {'random_seed': 11, 't_final': 2.0e-6, 'n_traj': 1,
                    'output_data': {'energy': {'write': 1}, 'delg_0': {'write': 1}}})


def doped_case(tag, species_count, extra):
    """Reference material_preprod + material_run with dopants (BVO, W on V sites): the doping
    hooks of do_kmc_steps (core.py:2723-2776): per-trajectory site_indices.npy -> shifted
    system_relative_energies, dopant charges in charge_config, site_charge_initiation."""
    core = rh.load_reference()
    from PyCD.material_preprod import material_preprod
    from PyCD.material_run import material_run
    work = rh.stage_example('BVO', TMP / tag, species_count=species_count, extra=extra)
    material_preprod(work)
    t0 = time.time()
    material_run(work)
    dt = time.time() - t0
    sim = yaml.safe_load(open(work / 'simulation_parameters.yml'))
    n_traj = int(sim['n_traj'])
    out = {'species_count': np.asarray(species_count), 'n_traj': n_traj, 'sim_yaml': yaml.safe_dump(sim),
           'example': 'BVO', 'ref_seconds': dt}
    core, run, _ = reference_run_object(work)
    total_steps = 0
    for i in range(n_traj):
        d = work / f'traj{i + 1}'
        out[f'rnd_state_{i}'] = np.frombuffer((d / 'initial_rnd_state.dump').read_bytes(), dtype=np.uint8)
        out[f'site_indices_{i}'] = np.load(d / 'site_indices.npy')
        out[f'unwrapped_{i}'] = np.load(d / 'unwrapped_traj.npy')
        td = np.load(d / 'time_data.npy')
        out[f'time_{i}'] = td
        total_steps += len(td) - 1
        # first-step state and rates from the reference's own routines, doping hooks applied
        # the way do_kmc_steps applies them (core.py:2723-2772)
        si = out[f'site_indices_{i}']
        dop = {}
        for row in si[si[:, 3] == 0]:
            dop.setdefault(run.dopant_element_types[row[1]], []).append(int(row[0]))
        e_rel = np.copy(run.undoped_system_relative_energies)
        for site, ti, _, shell in si:
            sub = run.substitution_element_types[ti]
            table = run.relative_energies['doping'][sub][ti]
            if shell < len(table):
                e_rel[site] += table[shell] * core.constants.EV2HARTREE
        run.system_relative_energies = e_rel
        core.rnd.setstate(pickle.load(open(d / 'initial_rnd_state.dump', 'rb')))
        occ = run.generate_initial_occupancy(dop)
        q = run.charge_config(occ, dop)
        attrs = run.get_process_attributes(occ)
        k_list, dg0, hopv = run.get_process_rates(attrs, q)
        out[f'occ0_{i}'] = np.asarray(occ)
        out[f'q0_{i}'] = np.asarray(q)[:, 0]
        out[f'e_rel_{i}'] = e_rel
        out[f'rates0_{i}'] = np.asarray(k_list)
        out[f'dg0_0_{i}'] = np.asarray(dg0)
    print(f'{tag}: {total_steps} steps in {dt:.1f} s (reference CPU, doped)')
    np.savez_compressed(GOLD / f'ref_{tag}.npz', **out)


def doped_cases():
    dope = {'num_dopants': [2, 0], 'min_shell_separation': [1, 1]}
    doped_case('bvo_doped_init', [2, 0],
               {'doping': dict(dope, site_charge_initiation=['yes', 'no']), 'random_seed': 3,
                't_final': 1.0e-4, 'n_traj': 2})
    doped_case('bvo_doped', [3, 0],
               {'doping': dict(dope, site_charge_initiation=['no', 'no']), 'random_seed': 4,
                't_final': 1.0e-4, 'n_traj': 2})


def main():
    if not rh.reference_available():
        raise SystemExit('reference not available; golden vectors are already committed')
    TMP.mkdir(exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'energy':
        energy_case()
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'doped':
        doped_cases()
        return
    copy_example('Hematite')
    copy_example('BVO')
    shipped_msd('Hematite')
    shipped_msd('BVO')
    field_on = {'external_field': {'electric': {'active': 1, 'dir': [1, 0, 0], 'ld': 0, 'mag': 0.002}}}
    reference_case('hematite_4e', 'Hematite', [4, 0],
                   {'random_seed': 7, 't_final': 1.0e-5, 'n_traj': 2, 'msd_t_final': 5000.0}, with_msd=True)
    reference_case('hematite_4e_field', 'Hematite', [4, 0],
                   dict(field_on, random_seed=7, t_final=2.0e-7, time_interval=1.0e-9, n_traj=2))
    reference_case('bvo_4e', 'BVO', [4, 0], {'random_seed': 2, 't_final': 3.0e-4, 'n_traj': 1})
    reference_case('bvo_2h', 'BVO', [0, 2], {'random_seed': 5, 't_final': 2.0e-6, 'n_traj': 1})
    energy_case()
    doped_cases()


if __name__ == '__main__':
    main()
