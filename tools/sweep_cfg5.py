#!/usr/bin/env python
"""BASELINE config 5: "Hematite temperature/field sweep, 8 conditions x 1024 trajectories
each, MSD/diffusivity reduction".

Conditions = T in {250, 300, 350, 400} K x field in {off, on: dir [1,0,0], mag 1e-4 a.u.}.
All conditions x trajectories are flattened into ONE trajectory list with per-trajectory
(kT, field) (pycd_kmc_ensemble_desc.kT_traj / field_traj); the list is split contiguously over
the GPUs, each condition's trajectories being interleaved so every GPU serves every condition.
Per condition: MSD (time-origin averaged), diffusivity +- SEM from per-trajectory slopes, drift
mobility; the per-trajectory arrays are gathered with one collective at the end.

    python tools/sweep_cfg5.py [--size 10 10 10] [--carriers 64] [--traj 1024] [--kmc-steps 20000]
    torchrun --nproc-per-node N tools/sweep_cfg5.py ...
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, nargs=3, default=[10, 10, 10])
    ap.add_argument('--carriers', type=int, default=64)
    ap.add_argument('--traj', type=int, default=1024, help='trajectories per condition')
    ap.add_argument('--kmc-steps', type=int, default=20000, help='KMC steps per trajectory')
    ap.add_argument('--n-path', type=int, default=1001)
    ap.add_argument('--n-msd', type=int, default=501)
    ap.add_argument('--trim', type=int, default=50)
    ap.add_argument('--refresh', type=int, default=64)
    ap.add_argument('--seed', type=int, default=2)
    args = ap.parse_args()
    import torch
    import yaml
    from pycd_b200 import _native as nat
    from pycd_b200 import constants
    from pycd_b200 import dist as D
    from pycd_b200 import ewald as EW
    from pycd_b200 import kmc as K
    from pycd_b200 import msd as M
    from pycd_b200.lattice import Lattice, Supercell
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
        os.dup2(saved, 1)
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(local)
    d = ROOT / 'tests' / 'golden' / 'hematite'
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    sim = yaml.safe_load(open(d / 'simulation_parameters.yml'))
    lat = Lattice(SimpleNamespace(**cfg))
    sc = Supercell(lat, args.size, [1, 1, 1])
    run = K.RunParameters(lat, sc, sc.hop_neighbor_tables(), 300, 'full', 'full', sim['t_final'],
                          sim['time_interval'], [args.carriers, 0], {}, sim['relative_energies'],
                          sim['external_field'])
    ep = EW.EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    t0 = time.perf_counter()
    p_unit = torch.empty((sc.n_per_cell, sc.num_system_elements), dtype=torch.float64, device=dev)
    EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell, out=p_unit.data_ptr())
    t_ewald = time.perf_counter() - t0
    system = K.KmcSystem(ctx, run, p_unit.data_ptr(), layout='unit_rows')

    temps = [250.0, 300.0, 350.0, 400.0]
    fields = [np.zeros(3), np.array([1e-4, 0.0, 0.0])]
    conds = [(T, f) for T in temps for f in fields]
    n_cond, n_total = len(conds), len(conds) * args.traj
    # global trajectory g -> condition g % n_cond (interleaved), replica g // n_cond
    lo, hi = D.block(rank, world, n_total)
    gids = np.arange(lo, hi)
    cond_of = gids % n_cond
    kT = np.array([conds[c][0] * constants.K2AUTEMP for c in cond_of])
    fld = np.array([conds[c][1] for c in cond_of])
    occ = K.philox_initial_occupancy(run.tables, hi - lo, args.carriers, args.seed, traj_id0=lo)
    S = args.kmc_steps - args.kmc_steps % args.refresh
    # grid so that the slowest condition (250 K) fills it: k ~ exp(-0.252 eV / kT)
    k250 = args.carriers * 3.2e9 * np.exp(-0.252 / 8.617e-5 * (1 / 250.0 - 1 / 300.0))
    dt_grid = (S / k250 * constants.SEC2AUTIME) / (args.n_path - 1)
    ens = K.KmcEnsemble(system, occ, dt_grid=dt_grid, n_path=args.n_path, step_limit=10 ** 12,
                        stop_at_grid_end=True, rng_mode=nat.RNG_PHILOX, seed=args.seed, traj_id0=lo,
                        refresh_interval=args.refresh, kT_traj=kT, field_traj=fld)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    chunk = 4096 - 4096 % args.refresh
    while ens.advance_resident(chunk) > 0:
        pass
    t_kmc = time.perf_counter() - t0
    state = ens.read(unwrapped=False)
    toff = np.array([0, args.carriers], dtype=np.int32)
    avg = M.species_avg_sd(ctx, ens.unwrapped_device_ptr(), hi - lo, args.n_path, args.carriers, args.n_msd,
                           1 / constants.ANG2BOHR, toff)
    drift = state['drift']
    steps = state['n_steps']
    if dist:
        avg = D.gather_trajectory_arrays(avg, n_total, device=dev)
        drift = D.gather_trajectory_arrays(drift, n_total, device=dev)
        steps = D.gather_trajectory_arrays(steps.astype(np.float64)[:, None], n_total, device=dev)[:, 0]
        tt = torch.tensor([t_kmc], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_kmc = float(tt[0])
    if rank == 0:
        all_cond = np.arange(n_total) % n_cond
        rows = []
        for c, (T, f) in enumerate(conds):
            sel = all_cond == c
            mp = SimpleNamespace(n_msd=args.n_msd, time_interval=dt_grid, time_conversion=constants.AUTIME2NS,
                                 trim_length=args.trim, n_dim=3, kBT=constants.KB * T / constants.EV2J)
            res = M.analyse(mp, avg[sel])
            row = {'T_K': T, 'field_au': float(f[0]), 'n_traj': int(sel.sum()),
                   'mean_steps': float(np.mean(steps[sel])),
                   'D_cm2_per_Vs': float(res['diffusivity'][0]), 'D_sem': float(res['diffusivity_sem'][0])}
            if f[0] != 0:
                mob = K.drift_mobility(drift[sel], f, float(np.linalg.norm(f))).mean(axis=1)
                row['drift_mobility_cm2_per_Vs'] = float(mob.mean())
                row['drift_mobility_sem'] = float(mob.std() / np.sqrt(len(mob)))
            rows.append(row)
        print(json.dumps({'config': f'Hematite {args.size} sweep, {n_cond} conditions x {args.traj} trajectories, '
                                    f'{args.carriers} electrons, {world} GPU(s)',
                          'kmc_seconds': round(t_kmc, 3), 'total_kmc_steps': float(np.sum(steps)),
                          'steps_per_second': float(np.sum(steps)) / t_kmc, 'ewald_seconds': round(t_ewald, 3),
                          'conditions': rows}))
    ens.close()
    system.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
