"""Drop-in for PyCD/material_run.py:13-157 + the file handling of Run.do_kmc_steps
(core.py:2647-2917): same YAML, same input files, same outputs; the step loop runs on
the GPU for all trajectories of the run at once.

Table layout.  Under full periodic boundaries material_setup also writes the rows of unit cell 0
(`precomputed_array_unit_rows.npy`, n_per_cell x N); material_run prefers them (`b200.p_layout: auto`):
the step kernel then reads the L2-resident lattice-stencil table (csrc/kmc_stencil.cuh) instead of
gathering from the N x N array, and a 10x10x10 Hematite run never touches a 7.2 GB file.

Several GPUs.  Launched under `torchrun` (RANK / WORLD_SIZE / LOCAL_RANK in the environment) the
trajectories of a serial run are split into contiguous blocks, one per rank / GPU; each rank writes its
own `traj<i>/`, rank 0 writes `drift_mobility.dat` and `Run.log` after the (kB-sized) gather of the drift
vectors.  Trajectory i uses the same random stream whatever the GPU count.
"""
import json
import os
from datetime import datetime

import numpy as np

from . import _native as nat
from . import dist as D
from . import ewald as ew
from . import kmc
from .config import input_directory, load_material_parameters, load_simulation_parameters
from .fileio import generate_report
from .lattice import Lattice, Supercell
from .material_setup import UNIT_ROWS_FILE
from .tables import load_hop_neighbor_list

DEFAULT_PHILOX_REFRESH = 256   # incremental rate updates, full re-gather every 256 steps


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


def rank_env():
    """(rank, world, local_rank) of a torchrun launch; (0, 1, 0) otherwise."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


class PhiloxSampler:
    """`random.Random`-like source for generate_initial_occupancy in Philox mode: one counter-based
    stream per GLOBAL trajectory id, so that sharding over GPUs is invisible.  With no explicit and no
    dopant-initiated sites, sample(pool, n) over the element's sites in ascending order equals
    kmc.philox_initial_occupancy."""

    def __init__(self, seed, traj_id):
        self.g = np.random.Generator(np.random.Philox(key=[int(seed), int(traj_id)]))

    def sample(self, pool, n):
        pool = list(pool)
        return [pool[i] for i in self.g.choice(len(pool), size=n, replace=False)]


def select_table(inp, supercell, opts):
    """(layout, array) of the Ewald table the run reads: b200.p_layout = auto | unit_rows | dense."""
    want = opts.get('p_layout', 'auto')
    if want not in ('auto', 'unit_rows', 'dense'):
        raise ValueError(f"b200.p_layout must be auto, unit_rows or dense, not {want!r}")
    unit_file, dense_file = inp / UNIT_ROWS_FILE, inp / 'precomputed_array.npy'
    full_pbc = ew.can_use_translation_symmetry(supercell)
    if want == 'unit_rows' or (want == 'auto' and full_pbc and unit_file.exists()):
        if not full_pbc:
            raise ValueError("b200.p_layout: unit_rows needs pbc = [1, 1, 1] and more than one cell")
        if not unit_file.exists():
            raise FileNotFoundError(f'{unit_file}: run material_setup(..., generate_precomputed_array=1)')
        return 'unit_rows', np.load(unit_file)
    return 'dense', np.load(dense_file)


def dense_array(ctx, supercell, layout, table, limit_gb=4.0):
    """The N x N array (energy outputs only: E0 = ewald_neut + q.P.q, core.py:2773-2780)."""
    if layout == 'dense':
        return table
    n = supercell.num_system_elements
    if n * n * 8 > limit_gb * 1e9:
        raise NotImplementedError('energy / delg_0 outputs need the dense array, which exceeds '
                                  f'{limit_gb} GB at N = {n}')
    return ew.ewald_expand(ctx, supercell, table, 0, n)


def advance_ensemble(ens, chunk, want_times=False, on_chunk=None):
    """Runs every trajectory of `ens` to its end (time grid full / step limit) in launches of `chunk`
    KMC steps.  Returns the per-trajectory event times (incl. the leading 0.0) when want_times."""
    pieces = [[np.zeros(1)] for _ in range(ens.n_traj)] if want_times else None
    while True:
        if want_times:
            res = ens.advance(chunk, want_times=True)
            for i in range(ens.n_traj):
                n = int(res['steps_done'][i])
                if n:
                    pieces[i].append(res['times'][i, :n].copy())
            active = res['n_active']
        else:
            active = ens.advance_resident(chunk)
        if on_chunk:
            on_chunk(active)
        if active == 0:
            break
    return [np.concatenate(p) for p in pieces] if want_times else None


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

    # trajectories of this rank (torchrun: contiguous blocks; `parallel` mode keeps the reference's
    # one-trajectory-per-process meaning and is never sharded)
    rank, world, local_rank = (0, 1, 0) if parallel else rank_env()
    lo, hi = D.block(rank, world, n_traj)
    n_local = hi - lo
    if not parallel:
        kmc.write_initial_rnd_states(dst_path, n_traj, sim['random_seed'], only=range(lo, hi))  # core.py:2655-2656
    traj_dirs = [dst_path if parallel else dst_path / f'traj{i + 1}' for i in range(lo, hi)]
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

    state, times = None, None
    info = {'rank': rank, 'world': world, 'trajectories': [lo, hi]}
    if n_local:
        # PYCD_B200_DEVICE pins the CUDA ordinal (several ranks sharing one GPU in tests)
        ctx = nat.default_context(int(os.environ.get('PYCD_B200_DEVICE', local_rank)))
        layout, table = select_table(inp, supercell, opts)
        system = kmc.KmcSystem(ctx, run, table, layout=layout)
        rng_kind = opts.get('rng', 'replay')
        chunk = int(opts.get('chunk_steps', 32768))
        want_times = bool(out_cfg['time']['write'])
        P_dense = dense_array(ctx, supercell, layout, table) if want_energy else None
        if rng_kind == 'replay':
            rngs = [kmc.load_rnd_state(d / 'initial_rnd_state.dump') for d in traj_dirs]
        elif rng_kind == 'philox':
            seed = int(sim['random_seed'])
            rngs = [PhiloxSampler(seed, lo + i) for i in range(n_local)]
        else:
            raise ValueError(f"b200.rng must be 'replay' or 'philox', not {rng_kind!r}")
        # generate_initial_occupancy (core.py:2478-2528): dopant-initiated, explicit, then sampled sites
        occ = np.array([run.initial_occupancy_from(r, traj_dopants(i)) for i, r in enumerate(rngs)],
                       dtype=np.int32).reshape(n_local, run.n_carriers)
        energy0 = np.array([run.initial_energy(P_dense, o, alpha_log, traj_q_lat(i)) for i, o in enumerate(occ)]) \
            if want_energy else None
        with nat.nvtx_range('pycd.kmc_run'):
            if rng_kind == 'replay':
                state, times, _ = kmc.run_replay(system, rngs, occ, chunk_steps=chunk, want_times=want_times,
                                                 energy0=energy0, doping=doping)
                info['step_kernel'] = state['last_kernel']
            else:
                refresh = int(opts.get('refresh_interval', DEFAULT_PHILOX_REFRESH))
                chunk = max(refresh, chunk - chunk % refresh)
                ens = kmc.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=lo,
                                      refresh_interval=refresh, energy0=energy0, doping=doping)
                times = advance_ensemble(ens, chunk, want_times=want_times)
                state = ens.read()
                info['step_kernel'] = ens.last_kernel()
                info['refresh_interval'] = refresh
                if want_energy:
                    state['energy_grid'], state['dg0_grid'] = ens.read_energy()
                ens.close()
        info.update(p_layout=layout, rng=rng_kind, kmc_steps=int(np.sum(state['n_steps'])),
                    near_tie_fallbacks=int(np.sum(state['near_tie'])), stencil=system.stencil_info()[0],
                    device=ctx.device)
        system.close()
    # sidecar of this implementation (not a reference file): which table / kernel / GPU served the rank
    with open(dst_path / (f'Run.b200.rank{rank}.json' if world > 1 else 'Run.b200.json'), 'w') as fh:
        json.dump(info, fh)

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

    drift = state['drift'] if state is not None else np.zeros((0, run.n_carriers, 3))
    if world > 1:
        drift = D.gather_blocks(drift, n_traj, local_rank=local_rank)
    if rank == 0:
        prefix = []
        if run.field_active:
            mob = kmc.drift_mobility(drift, run.field, run.field_mag)
            np.savetxt(dst_path / 'drift_mobility.dat', mob)
            sp_mean = mob.mean(axis=1)  # one carrier type: all carriers (core.py:2066-2081)
            prefix.append(f'Estimated value of {run.species_type} drift mobility is: '
                          f'{np.mean(sp_mean):4.3e} cm2/V.s.\n')
            prefix.append(f'Standard error of mean in {run.species_type} drift mobility is: '
                          f'{np.std(sp_mean) / np.sqrt(n_traj):4.3e} cm2/V.s.\n')
        generate_report(start_time, dst_path, 'Run', 1, ''.join(prefix))
    if world > 1:
        D.barrier()
    return None
