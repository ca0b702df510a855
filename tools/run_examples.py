#!/usr/bin/env python
"""BASELINE configs 1 and 2 through the drop-in drivers, timed: the reference's shipped
Hematite example (1 electron, 319 946 KMC steps) and the BVO example inputs with 4 electrons
(fixed 100 000 steps, replayed MT19937 draws), plus Ewald setup and MSD.  Prints one JSON line.
These cases are latency-bound (one trajectory, 4-24 processes): steps/s only, no roofline."""
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'tests'):
    sys.path.insert(0, str(p))


def main():
    import helpers as H
    from pycd_b200 import _native as nat
    from pycd_b200 import kmc as K
    from pycd_b200 import material_msd, material_run, material_setup
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(0)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        # ---- config 1: examples/Hematite as shipped ----
        work = tmp / 'hematite'
        shutil.copytree(H.GOLD / 'hematite', work)
        shutil.rmtree(work / 'traj1')
        ex = H.load_example('hematite')
        t0 = time.perf_counter()
        material_setup(work / 'InputFiles', np.array([2, 2, 1]), np.array([1, 1, 1]), 1, 0, 1, 0, 0)
        t_setup = time.perf_counter() - t0
        np.save(work / 'InputFiles' / 'precomputed_array.npy', ex.P)  # shipped array -> shipped trajectory
        shutil.copy(H.GOLD / 'hematite' / 'InputFiles' / 'hop_neighbor_list.npy', work / 'InputFiles')
        t0 = time.perf_counter()
        material_run(work)
        t_run = time.perf_counter() - t0
        n_steps = len(np.load(work / 'traj1' / 'time_data.npy')) - 1
        same = bool(np.array_equal(np.load(work / 'traj1' / 'unwrapped_traj.npy'), H.shipped_unwrapped(ex)))
        t0 = time.perf_counter()
        material_msd(work)
        t_msd = time.perf_counter() - t0
        out['config1_hematite_shipped'] = {
            'kmc_steps': n_steps, 'material_run_seconds': round(t_run, 3),
            'steps_per_second_incl_io': n_steps / t_run, 'kernel_ms': ctx.total_kernel_ms(nat.KC_KMC_STEP),
            'steps_per_second_kernel': n_steps / (ctx.total_kernel_ms(nat.KC_KMC_STEP) * 1e-3),
            'trajectory_identical_to_shipped': same, 'setup_seconds_neighbours_plus_ewald': round(t_setup, 3),
            'material_msd_seconds': round(t_msd, 3),
            'reference_cpu': '15 489 steps/s (20.66 s), Ewald 1.38 s, MSD 0.31 s (BASELINE.md, survey container)'}
        # ---- config 2: BVO inputs, 4 electrons, fixed 100 000 steps, replay ----
        ctx.reset_timers()
        exb = H.load_example('bvo', species_count=[4, 0])
        run = H.run_parameters(exb)
        import random
        rng = random.Random(2)
        occ = run.initial_occupancy_from(rng)
        system = K.KmcSystem(ctx, run, exb.P)
        steps = 100000
        draws = np.array([rng.random() for _ in range(2 * steps)])
        ens = K.KmcEnsemble(system, np.array([occ]), rng_mode=nat.RNG_REPLAY, step_limit=steps,
                            stop_at_grid_end=False)
        t0 = time.perf_counter()
        ens.advance(steps, draws=draws[None, :])
        t_run = time.perf_counter() - t0
        got = ens.read(unwrapped=False)
        out['config2_bvo_4e_100k_steps'] = {
            'kmc_steps': int(got['n_steps'][0]), 'advance_seconds': round(t_run, 3),
            'steps_per_second': int(got['n_steps'][0]) / t_run,
            'steps_per_second_kernel': int(got['n_steps'][0]) / (ctx.total_kernel_ms(nat.KC_KMC_STEP) * 1e-3),
            'n_proc': run.n_proc, 'reference_cpu': '~4 300 steps/s (BVO 4 e-, make_golden.py run in the build container)'}
        ens.close()
        system.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
