"""Drop-in for PyCD/material_run.py:13-157 + the file handling of Run.do_kmc_steps
(core.py:2647-2917): same YAML, same input files, same outputs; the step loop runs on
the GPU for all trajectories of the run at once."""
from datetime import datetime

import numpy as np

from . import _native as nat
from . import kmc
from .config import input_directory, load_material_parameters, load_simulation_parameters
from .fileio import generate_report
from .lattice import Lattice, Supercell
from .tables import load_hop_neighbor_list


def _alpha_from_log(path, fallback):
    """material_run.py:75-80: alpha is re-read from line 3 of precomputed_array.log."""
    try:
        with open(path, 'r') as fh:
            for i, line in enumerate(fh):
                if i == 3:
                    return float(line[7:16])
    except OSError:
        pass
    return fallback


def material_run(dst_path):
    start_time = datetime.now()
    sim = load_simulation_parameters(dst_path)
    inp = input_directory(dst_path, sim)
    params = load_material_parameters(inp)
    lattice = Lattice(params)
    supercell = Supercell(lattice, sim['system_size'], sim['pbc'])
    if (dst_path / 'Run.log').exists() and not sim['over_write']:
        print('Simulation files already exists in the destination directory')
        return None
    opts = sim['b200']
    hop = load_hop_neighbor_list(inp / 'hop_neighbor_list.npy')
    alpha_log = _alpha_from_log(inp / 'precomputed_array.log', params.alpha)  # feeds the energy output only
    P = np.load(inp / 'precomputed_array.npy')
    parallel = sim['compute_mode'] == 'parallel'
    n_traj = 1 if parallel else int(sim['n_traj'])
    run = kmc.RunParameters(lattice, supercell, hop, sim['temp'], sim['ion_charge_type'],
                            sim['species_charge_type'], sim['t_final'], sim['time_interval'],
                            sim['species_count'], sim['initial_occupancy'],
                            sim['relative_energies'], sim['external_field'], sim.get('doping'))
    out_cfg = sim['output_data']
    want_energy = bool(out_cfg.get('energy', {}).get('write') or out_cfg.get('delg_0', {}).get('write'))
    if out_cfg['unwrapped_traj'].get('write_every_step'):
        raise NotImplementedError('write_every_step is not supported (SURVEY appendix D)')

    if not parallel:
        kmc.write_initial_rnd_states(dst_path, n_traj, sim['random_seed'])  # core.py:2655-2656
    traj_dirs = [dst_path if parallel else dst_path / f'traj{i + 1}' for i in range(n_traj)]
    for d in traj_dirs:
        d.mkdir(parents=True, exist_ok=True)

    # Doping hooks (core.py:2723-2776): every trajectory reads the site_indices.npy that the
    # reference's material_preprod wrote (the dopant distribution itself is control plane and not
    # regenerated here, unlike the reference's serial mode, which rewrites the same files).
    doping = None
    if run.doping.active:
        doping = []
        for d in traj_dirs:
            path = run.doping.site_indices_path(d)
            if not path.exists():
                raise FileNotFoundError(f'{path}: doped runs need the site_indices.npy written by '
                                        "PyCD's material_preprod (dopant distribution, core.py:2566-2645)")
            doping.append(run.doping.load(np.load(path), run.e_rel, run.q_lat))

    def traj_q_lat(i):
        return doping[i].q_lat(run.q_lat) if doping else None

    def traj_dopants(i):
        return doping[i].dopant_site_indices if doping else None

    ctx = nat.default_context()
    system = kmc.KmcSystem(ctx, run, P)
    rng_kind = opts.get('rng', 'replay')
    chunk = int(opts.get('chunk_steps', 32768))
    want_times = bool(out_cfg['time']['write'])
    if rng_kind == 'replay':
        rngs = [kmc.load_rnd_state(d / 'initial_rnd_state.dump') for d in traj_dirs]
        occ = np.array([run.initial_occupancy_from(r, traj_dopants(i)) for i, r in enumerate(rngs)],
                       dtype=np.int32)
        energy0 = np.array([run.initial_energy(P, o, alpha_log, traj_q_lat(i)) for i, o in enumerate(occ)]) \
            if want_energy else None
        state, times, _ = kmc.run_replay(system, rngs, occ, chunk_steps=chunk, want_times=want_times,
                                         energy0=energy0, doping=doping)
    elif rng_kind == 'philox':
        seed = int(sim['random_seed'])
        occ = kmc.philox_initial_occupancy(run.tables, n_traj, run.n_carriers, seed)
        refresh = int(opts.get('refresh_interval', 1))
        chunk -= chunk % refresh
        energy0 = np.array([run.initial_energy(P, o, alpha_log, traj_q_lat(i)) for i, o in enumerate(occ)]) \
            if want_energy else None
        ens = kmc.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=seed, refresh_interval=refresh,
                              energy0=energy0, doping=doping)
        pieces = [[np.zeros(1)] for _ in range(n_traj)]
        while True:
            res = ens.advance(chunk, want_times=want_times)
            if want_times:
                for i in range(n_traj):
                    n = int(res['steps_done'][i])
                    if n:
                        pieces[i].append(res['times'][i, :n].copy())
            if res['n_active'] == 0:
                break
        state = ens.read()
        if want_energy:
            state['energy_grid'], state['dg0_grid'] = ens.read_energy()
        ens.close()
        times = [np.concatenate(p) for p in pieces] if want_times else None
    else:
        raise ValueError(f"b200.rng must be 'replay' or 'philox', not {rng_kind!r}")
    system.close()

    n_path, c3 = run.n_path, 3 * run.n_carriers
    for i, d in enumerate(traj_dirs):
        for kind, attrs in out_cfg.items():
            if not attrs.get('write'):
                continue
            target = d / attrs['file_name']
            if kind == 'unwrapped_traj':
                np.save(target, state['unwrapped'][i])
            elif kind == 'time':
                np.save(target, times[i])
            elif kind == 'energy':
                np.save(target, state['energy_grid'][i])
            elif kind == 'delg_0':
                np.save(target, state['dg0_grid'][i])
            elif kind == 'wrapped_traj':  # allocated, never filled by the reference (core.py:2712-2714)
                np.save(target, np.zeros((n_path, c3)))
            elif kind == 'potential':     # likewise (core.py:2719-2721)
                np.save(target, np.zeros((n_path, run.n_carriers)))
        h5 = out_cfg.get('hdf5_output', {})
        if h5.get('enabled') and out_cfg['unwrapped_traj'].get('write'):
            from .hdf5_io import write_trajectory_h5
            write_trajectory_h5(d / h5['file_name'], state['unwrapped'][i], run.n_carriers,
                                run.time_interval)

    prefix = []
    if run.field_active:
        mob = kmc.drift_mobility(state['drift'], run.field, run.field_mag)
        np.savetxt(dst_path / 'drift_mobility.dat', mob)
        sp_mean = mob.mean(axis=1)  # one carrier type: all carriers (core.py:2066-2081)
        prefix.append(f'Estimated value of {run.species_type} drift mobility is: '
                      f'{np.mean(sp_mean):4.3e} cm2/V.s.\n')
        prefix.append(f'Standard error of mean in {run.species_type} drift mobility is: '
                      f'{np.std(sp_mean) / np.sqrt(n_traj):4.3e} cm2/V.s.\n')
    generate_report(start_time, dst_path, 'Run', 1, ''.join(prefix))
    return None
