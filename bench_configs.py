"""The BASELINE.json configurations other than the headline one, as bounded legs of bench.py.

bench.py's headline line is config 3 (Hematite 10x10x10, 64 electrons, 512 trajectories per GPU).  The
functions below measure the other configurations on the same ranks, each checked against the CPU oracle
or a size-independent property inside the leg, and bench.py attaches their results to its JSON line
(`configs`), so that the driver's 1/2/4/8-GPU runs see them too:

  cfg2  BVO 4x4x2 (N = 768), 16 electrons, 8 neighbour slots per V (performance variant of
        "examples/BVO: multiple charge carriers"), lattice-stencil kernel, Philox ensemble
  cfg4  BVO 12x12x6 (N = 20 736, K_eff = 336 074) dense Ewald array: rows sharded over the ranks, evaluated
        directly (no translation symmetry), NCCL all-gather; STRONG scaling (the array is fixed)
  cfg5  Hematite 10x10x10 temperature / field sweep, 8 conditions flattened into one ensemble with
        per-trajectory (kT, field, time grid), MSD / diffusivity / drift mobility per condition; 1024
        trajectories per GPU in the default line (weak), `--config 5` runs the full 8 x 1024 (strong)

`python bench.py --config N` prints the leg as its own JSON line at full size.
"""
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent
GOLD = ROOT / 'tests' / 'golden'


def _lattice(example):
    import yaml
    from pycd_b200.lattice import Lattice
    d = GOLD / example
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    sim = yaml.safe_load(open(d / 'simulation_parameters.yml'))
    return Lattice(SimpleNamespace(**cfg)), cfg, sim


def _sync(dist):
    import torch
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()


def _max_over_ranks(dist, dev, *vals):
    if not dist:
        return vals
    import torch
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return tuple(float(v) for v in t)


# -----------------------------------------------------------------------------------------------
def cfg2_bvo(ctx, dev, rank, world, dist, traj_per_gpu=512, kmc_steps=8192, launches=4, refresh=256,
             size=(4, 4, 2), carriers=16, check_traj=8):
    """BVO electrons, 8 neighbour slots: steps/s of the stencil kernel + oracle check of the first
    trajectories of the timed ensemble (final sites and displacement grids, every step of every launch)."""
    import torch
    import oracle as O
    from pycd_b200 import _native as nat, constants, ewald as EW, kmc as K
    from pycd_b200.lattice import Supercell
    lat, cfg, sim = _lattice('bvo')
    sc = Supercell(lat, list(size), [1, 1, 1])
    run = K.RunParameters(lat, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full', sim['t_final'],
                          sim['time_interval'], [carriers, 0], {}, sim['relative_energies'], sim['external_field'])
    ep = EW.EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    n = sc.num_system_elements
    coords = np.ascontiguousarray(sc.coordinates)
    p_unit, _ = EW.unit_cell_rows(ctx, ep, coords)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    seed, n_path = 11, 101
    S = kmc_steps - kmc_steps % refresh
    occ = K.philox_initial_occupancy(run.tables, traj_per_gpu, carriers, seed, traj_id0=rank * traj_per_gpu)
    # BVO V:V hops are ~1e3 slower than Hematite's; the grid only has to fill during the run
    probe = K.KmcEnsemble(system, occ[:4], dt_grid=1e30, n_path=2, step_limit=256, stop_at_grid_end=False,
                          rng_mode=nat.RNG_PHILOX, seed=seed, refresh_interval=1)
    probe.advance_resident(256)
    t_probe = float(probe.read(unwrapped=False)['time'].mean())
    probe.close()
    dt_grid = t_probe / 256 * S * (launches + 1) / (n_path - 1)
    ens = K.KmcEnsemble(system, occ, dt_grid=dt_grid, n_path=n_path, step_limit=10 ** 12, stop_at_grid_end=False,
                        rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=rank * traj_per_gpu, refresh_interval=refresh)
    ens.advance_resident(S)
    _sync(dist)
    ctx.reset_timers()
    t0 = time.perf_counter()
    for _ in range(launches):
        ens.advance_resident(S)
    _sync(dist)
    wall = time.perf_counter() - t0
    kern_ms = ctx.total_kernel_ms(nat.KC_KMC_STEP) / launches
    kernel = ens.last_kernel()
    state = ens.read(unwrapped=True)
    ens.close()
    wall, kern_ms = _max_over_ranks(dist, dev, wall, kern_ms)
    out = {'workload': f'BVO {size[0]}x{size[1]}x{size[2]} (N={n}), {carriers} electrons, 8 neighbour slots, '
                       f'{traj_per_gpu} trajectories/GPU, Philox, refresh {refresh}',
           'value': world * traj_per_gpu * S * launches / wall, 'unit': 'KMC steps/s', 'kernel': kernel,
           'kernel_ms_per_launch': kern_ms, 'kmc_steps_per_launch': S, 'launches': launches}
    if rank == 0:
        dense = EW.ewald_expand(ctx, sc, p_unit, 0, n)
        steps = int(state['n_steps'][0])
        orc = O.KmcOracle(run, dense, dt_grid=dt_grid, n_path=n_path, step_limit=steps, stop_at_grid_end=False,
                          rng_mode=1, seed=seed)
        ref = orc.ensemble(occ[:check_traj], traj_id0=0)
        same = [bool(np.array_equal(ref['occupancy'][i], state['occupancy'][i]) and
                     np.array_equal(ref['unwrapped'][i], state['unwrapped'][i])) for i in range(check_traj)]
        out['parity'] = {'checked': check_traj, 'equal': int(sum(same)), 'kmc_steps_per_trajectory': steps,
                         'what': 'final sites and displacement grids vs the CPU oracle on the expanded array'}
    system.close()
    return out


# -----------------------------------------------------------------------------------------------
def cfg4_ewald(ctx, dev, rank, world, dist, size=(12, 12, 6), check_rows=64, example='bvo'):
    """Dense Ewald array of a large BVO supercell: rank g evaluates rows [gN/G, (g+1)N/G) directly, one
    NCCL all-gather completes the array on every GPU.  Checked (i) on `check_rows` rows of every rank's block
    against the translation-expanded unit-cell rows (an independent evaluation path: different tile shape,
    different split-k) and (ii) for symmetry on the gathered array."""
    import torch
    from pycd_b200 import dist as D, ewald as EW
    from pycd_b200.lattice import Supercell
    lat, cfg, _ = _lattice(example)
    sc = Supercell(lat, list(size), [1, 1, 1])
    ep = EW.EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    n = sc.num_system_elements
    lo, hi = D.block(rank, world, n)
    P = torch.empty((n, n), dtype=torch.float64, device=dev)
    coords = torch.from_numpy(np.ascontiguousarray(sc.coordinates)).to(dev)
    _sync(dist)
    t0 = time.perf_counter()
    _, st = EW.ewald_rows(ctx, ep, coords.data_ptr(), lo, hi, out=P[lo:hi].data_ptr(), plan_rows=n)
    torch.cuda.synchronize()
    t_rows = time.perf_counter() - t0
    t_gather = 0.0
    if dist:
        tg = time.perf_counter()
        D.allgather_rows(P)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - tg
    total = time.perf_counter() - t0
    # checks (untimed)
    pu = torch.empty((sc.n_per_cell, n), dtype=torch.float64, device=dev)
    _, st_u = EW.unit_cell_rows(ctx, ep, coords.data_ptr(), out=pu.data_ptr())   # class-factorised sum: independent path
    r0 = lo + (hi - lo - check_rows) // 2
    ref = torch.empty((check_rows, n), dtype=torch.float64, device=dev)
    EW.ewald_expand(ctx, sc, pu.data_ptr(), r0, r0 + check_rows, out=ref.data_ptr())
    torch.cuda.synchronize()
    scale = float(pu.abs().max())
    err = float((P[r0:r0 + check_rows] - ref).abs().max())
    asym = float((P[r0:r0 + check_rows] - P[:, r0:r0 + check_rows].T).abs().max())
    flops = 4.0 * (hi - lo) * n * st['k_eff']
    tfl = flops / (st['fourier_ms'] * 1e-3) / 1e12
    total, t_rows, t_gather, err, asym, f_ms = _max_over_ranks(dist, dev, total, t_rows, t_gather, err, asym,
                                                               st['fourier_ms'])
    del P, pu, ref
    torch.cuda.empty_cache()
    return {'workload': f'{example.upper()} {size[0]}x{size[1]}x{size[2]} (N={n}) dense Ewald array, rows sharded over '
                        f'{world} GPU(s) + all-gather', 'value': total, 'unit': 's', 'scaling': 'strong',
            'n_sites': n, 'k_eff': int(st['k_eff']), 'rows_per_gpu': hi - lo, 'k_split': int(st['k_split']),
            'seconds_rows': t_rows, 'seconds_fourier_kernel': f_ms / 1e3, 'seconds_allgather': t_gather,
            'fp64_tflops_per_gpu': tfl, 'fp64_peak_tflops': 37.1, 'roofline_frac': tfl / 37.1,
            'symmetric_path': {'unit_rows_method': st_u.get('method'), 'unit_rows_fourier_ms': st_u['fourier_ms'],
                               'note': 'the same array from the class-factorised unit-cell rows + translation '
                                       'expansion (what material_setup does under full PBC); the dense evaluation '
                                       'above is the general path and the FP64 roofline kernel'},
            'parity': {'rows_checked_per_rank': check_rows, 'max_abs_diff_vs_translation_expanded': err,
                       'max_abs_P': scale, 'rel': err / scale, 'max_asymmetry': asym,
                       'ok': bool(err <= 1e-12 * scale and asym <= 1e-12 * scale)}}


# -----------------------------------------------------------------------------------------------
def cfg5_sweep(ctx, dev, rank, world, dist, system, run, traj_per_condition, kmc_steps=20480, n_path=1001,
               n_msd=501, trim=50, refresh=64, common_grid=False, chunk_steps=4096):
    """The temperature / field sweep on an existing KMC system (the headline system: Hematite 10x10x10)."""
    from pycd_b200 import sweep as SW
    conds = SW.hematite_conditions()
    intervals = SW.balanced_intervals(conds, run.n_carriers, kmc_steps, n_path)
    if common_grid:   # one grid for every condition: the slowest condition's (round-1 behaviour)
        intervals = np.full(len(conds), intervals.max())
    _sync(dist)
    t0 = time.perf_counter()
    res = SW.run_sweep(system, run, conds, traj_per_condition, intervals, n_path, seed=2, refresh_interval=refresh,
                       chunk_steps=chunk_steps, rank=rank, world=world, n_msd=n_msd, trim=trim, comm_device=dev)
    _sync(dist)
    wall = time.perf_counter() - t0
    wall, t_kmc = _max_over_ranks(dist, dev, wall, res['kmc_seconds'])
    total_steps = float(np.sum(res['n_steps']))
    return {'workload': f'Hematite sweep, {len(conds)} conditions x {traj_per_condition} trajectories '
                        f'({res["n_total"]} in all, {res["n_total"] // world} per GPU), {run.n_carriers} electrons, '
                        f'{"one common" if common_grid else "per-condition"} time grid of {n_path} rows',
            'value': total_steps / t_kmc, 'unit': 'KMC steps/s', 'seconds_kmc': t_kmc, 'seconds_total': wall,
            'total_kmc_steps': total_steps, 'kernel': res['kernel'], 'launches_per_rank': res['launches'],
            'near_tie_fallbacks': res['near_tie'], 'conditions': res['conditions']}
