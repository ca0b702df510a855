"""A temperature / field sweep as ONE ensemble (BASELINE config 5: "Hematite temperature/field sweep,
8 conditions x 1024 trajectories each, MSD/diffusivity reduction").

In the reference a sweep is one `material_run` + `material_msd` per condition, each with its own
simulation_parameters.yml (temp, external_field, t_final, time_interval: PyCD/material_run.py:15-27,
142-150).  Here all conditions x trajectories are flattened into one trajectory list with per-trajectory
(kT, field, time_interval) -- pycd_kmc_ensemble_desc.kT_traj / field_traj / dt_grid_traj -- so that one
launch keeps every SM busy with all conditions at once:

* global trajectory g serves condition g % n_cond (replica g // n_cond): every contiguous block of the
  list -- i.e. every GPU's shard -- holds all conditions in equal numbers, so the shards carry the same
  load whatever the condition costs;
* every condition keeps its own time grid (`time_interval[c]`, `n_path` rows): with intervals chosen in
  proportion to 1 / k_total(T) the trajectories of all conditions take about the same number of KMC steps
  and finish together (a grid common to all conditions makes the hot conditions run ~100x longer than the
  cold ones at 250-400 K and leaves most SMs idle for most of the run);
* per condition: MSD (csrc/msd.cu, Analysis.compute_msd core.py:2974-3081), diffusivity +- SEM from the
  per-trajectory slopes, drift mobility (core.py:2052-2082); the per-trajectory arrays are gathered with
  one collective at the end (dist.gather_trajectory_arrays).
"""
from types import SimpleNamespace

import numpy as np

from . import _native as nat
from . import constants
from . import dist as D
from . import kmc as K
from . import msd as M


def hematite_conditions(temps=(250.0, 300.0, 350.0, 400.0), field_mag=1e-4):
    """T x field grid of BASELINE config 5: field off / on along [1, 0, 0] (a.u.)."""
    return [(float(T), np.array(f, dtype=float)) for T in temps
            for f in ([0.0, 0.0, 0.0], [field_mag, 0.0, 0.0])]


def balanced_intervals(conditions, n_carriers, kmc_steps, n_path, rate_300=3.2e9, barrier_ev=0.252):
    """time_interval (a.u.) per condition such that a trajectory fills its n_path-row grid in about
    kmc_steps KMC steps: k_total ~ C * rate_300 * exp(-E_a/k_B (1/T - 1/300)) hops per second (Hematite
    electrons, basal-plane hop: lambda/4 - V_AB = 0.252 eV)."""
    out = []
    for T, _ in conditions:
        k_tot = n_carriers * rate_300 * np.exp(-barrier_ev / 8.617333262e-5 * (1.0 / T - 1.0 / 300.0))
        out.append(kmc_steps / k_tot * constants.SEC2AUTIME / (n_path - 1))
    return np.array(out)


def run_sweep(system, run, conditions, traj_per_condition, intervals, n_path, seed=2, refresh_interval=64,
              chunk_steps=4096, rank=0, world=1, n_msd=None, trim=None, comm_device=None, max_steps=10 ** 12):
    """Runs the rank's shard of the flattened sweep and reduces per condition.

    conditions: [(T_kelvin, field[3] a.u.)]; intervals: time_interval (a.u.) per condition.
    Returns a dict: per-condition rows (rank-independent), raw per-trajectory arrays of the WHOLE sweep
    (after the gather), timing of the KMC phase of this rank and the kernel that ran."""
    import time
    ctx = system.ctx
    n_cond = len(conditions)
    n_total = n_cond * int(traj_per_condition)
    C_ = run.n_carriers
    lo, hi = D.block(rank, world, n_total)
    gids = np.arange(lo, hi)
    cond_of = gids % n_cond
    kT = np.array([conditions[c][0] * constants.K2AUTEMP for c in cond_of])
    fld = np.array([conditions[c][1] for c in cond_of]).reshape(-1, 3)
    dtg = np.asarray(intervals, dtype=float)[cond_of]
    occ = K.philox_initial_occupancy(run.tables, hi - lo, C_, seed, traj_id0=lo)
    chunk = int(chunk_steps) - int(chunk_steps) % max(int(refresh_interval), 1)
    ens = K.KmcEnsemble(system, occ, dt_grid=float(dtg[0]), n_path=n_path, step_limit=max_steps,
                        stop_at_grid_end=True, rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=lo,
                        refresh_interval=refresh_interval, kT_traj=kT, field_traj=fld, dt_grid_traj=dtg)
    t0 = time.perf_counter()
    launches = 0
    with nat.nvtx_range('pycd.sweep_kmc'):
        while ens.advance_resident(chunk) > 0:
            launches += 1
    launches += 1
    t_kmc = time.perf_counter() - t0
    kernel = ens.last_kernel()
    state = ens.read(unwrapped=False)
    n_msd = int(n_msd or (n_path // 2 + 1))
    trim = int(trim if trim is not None else max(1, n_msd // 10))
    toff = np.array([0, C_], dtype=np.int32)
    with nat.nvtx_range('pycd.sweep_msd'):
        avg = M.species_avg_sd(ctx, ens.unwrapped_device_ptr(), hi - lo, n_path, C_, n_msd,
                               1 / constants.ANG2BOHR, toff)
    drift, steps, near = state['drift'], state['n_steps'].astype(np.float64), state['near_tie'].astype(np.float64)
    ens.close()
    if world > 1:
        avg = D.gather_trajectory_arrays(avg, n_total, device=comm_device)
        drift = D.gather_trajectory_arrays(drift, n_total, device=comm_device)
        both = D.gather_trajectory_arrays(np.stack([steps, near], axis=1), n_total, device=comm_device)
        steps, near = both[:, 0], both[:, 1]
    rows = analyse_conditions(conditions, intervals, avg, drift, steps, n_msd, trim)
    return {'conditions': rows, 'avg_sd': avg, 'drift': drift, 'n_steps': steps, 'near_tie': int(near.sum()),
            'kmc_seconds': t_kmc, 'launches': launches, 'kernel': kernel, 'n_total': n_total,
            'local_steps': float(state['n_steps'].sum()), 'n_msd': n_msd, 'trim': trim}


def analyse_conditions(conditions, intervals, avg, drift, steps, n_msd, trim):
    """Per-condition MSD slope -> diffusivity +- SEM (core.py:3052-3071) and drift mobility (core.py:2052-2082)
    from the whole sweep's per-trajectory arrays (trajectory g belongs to condition g % n_cond)."""
    n_cond = len(conditions)
    all_cond = np.arange(avg.shape[0]) % n_cond
    rows = []
    for c, (T, f) in enumerate(conditions):
        sel = all_cond == c
        mp = SimpleNamespace(n_msd=n_msd, time_interval=float(intervals[c]), time_conversion=constants.AUTIME2NS,
                             trim_length=trim, n_dim=3, kBT=constants.KB * T / constants.EV2J)
        res = M.analyse(mp, avg[sel])
        row = {'T_K': T, 'field_au': [float(v) for v in f], 'n_traj': int(sel.sum()),
               'time_interval_ns': float(intervals[c]) * constants.AUTIME2NS,
               'mean_steps': float(np.mean(steps[sel])),
               'D_cm2_per_Vs': float(res['diffusivity'][0]), 'D_sem': float(res['diffusivity_sem'][0]),
               'msd_last_A2': float(res['msd_data'][-1, 1])}
        mag = float(np.linalg.norm(f))
        if mag != 0:
            mob = K.drift_mobility(drift[sel], f, mag).mean(axis=1)
            row['drift_mobility_cm2_per_Vs'] = float(mob.mean())
            row['drift_mobility_sem'] = float(mob.std() / np.sqrt(len(mob)))
        rows.append(row)
    return rows
