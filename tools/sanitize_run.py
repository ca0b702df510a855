#!/usr/bin/env python
"""Small invocations of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py     # shared-memory hazards
    compute-sanitizer --tool synccheck python tools/sanitize_run.py     # barrier / warp-sync misuse

Covers: Ewald (skinny and dense tiles, split-k, finish, expansion), the lattice-stencil step kernels
(two warps / one warp, 4 and 8 neighbour slots, stateless / incremental, plain / featured with field and
energy outputs), the gather step kernels (dense and unit rows), V_lat and MSD.  Sizes are tiny: the
sanitizer serialises and instruments every access.  Every result is also compared with the CPU oracle,
so a sanitizer-clean run that computes garbage still fails."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'oracle', ROOT / 'tests'):
    sys.path.insert(0, str(p))


def main():
    import helpers as H
    import oracle as O
    from pycd_b200 import _native as nat, constants, ewald as EW, kmc as K, msd as M
    from pycd_b200.lattice import Supercell
    ctx = nat.default_context(0)
    done = []

    # ---- Ewald: shipped 2x2x1 cell (skinny tile + split-k via the symmetric path, dense tile directly)
    ex = H.load_example('hematite')
    ep = H.ewald_parameters(ex)
    coords = H.shipped_coords(ex)
    P, _ = EW.precomputed_array(ctx, ep, coords=coords, symmetric=False)
    assert np.abs(P - ex.P).max() <= 1e-10 * np.abs(ex.P).max()
    done.append('ewald dense rows (64x128 tile) + finish')
    sc = Supercell(ex.lattice, [2, 2, 2], [1, 1, 1])
    ep2 = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    c2 = np.ascontiguousarray(sc.coordinates)
    p_unit, st = EW.ewald_rows(ctx, ep2, c2, 0, sc.n_per_cell)
    dense = EW.ewald_expand(ctx, sc, p_unit, 0, sc.num_system_elements)
    direct, _ = EW.ewald_rows(ctx, ep2, c2, 0, sc.num_system_elements)
    assert np.abs(direct - dense).max() <= 1e-12 * np.abs(dense).max()
    parts = sum(EW.ewald_rows(ctx, ep2, c2, 0, sc.n_per_cell, k_part=p, k_parts=3)[0] for p in range(3))
    assert np.abs(parts - p_unit).max() <= 1e-13 * np.abs(p_unit).max()
    done.append(f'ewald unit-cell rows (32x256 tile, split-k {st["k_split"]}), expansion, k-range parts')

    # ---- stencil step kernels on Hematite 2x2x2, 40 electrons (two warps) and 20 (one warp)
    sim = ex.sim
    for carriers, kernel in ((40, 'kmc_step_warp_kernel<2,1,4>'), (20, 'kmc_step_warp_kernel<1,1,4>')):
        run = K.RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full', sim['t_final'],
                              sim['time_interval'], [carriers, 0], {}, sim['relative_energies'], sim['external_field'])
        occ = K.philox_initial_occupancy(run.tables, 3, carriers, seed=3)
        system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
        for refresh in (1, 16):
            kw = dict(dt_grid=run.time_interval / 300, n_path=16, step_limit=96, stop_at_grid_end=False)
            ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=3, refresh_interval=refresh, **kw)
            ens.advance_resident(96)
            got = ens.read()
            assert ens.last_kernel() == kernel, ens.last_kernel()
            ens.close()
            ref = O.KmcOracle(run, dense, rng_mode=1, seed=3, **kw).ensemble(occ)
            assert np.array_equal(got['occupancy'], ref['occupancy']) and np.array_equal(got['unwrapped'], ref['unwrapped'])
            done.append(f'{kernel} refresh {refresh} (plain)')
        # featured variant: field + per-step event / time outputs
        fld = np.array([[1e-4, 0.0, 0.0]] * 3)
        ens = K.KmcEnsemble(system, occ, rng_mode=nat.RNG_PHILOX, seed=3, refresh_interval=16, field_traj=fld,
                            dt_grid=run.time_interval / 300, n_path=16, step_limit=64, stop_at_grid_end=False)
        res = ens.advance(64, want_events=True, want_times=True)
        ens.close()
        one = O.KmcOracle(run, dense, field=fld[0], rng_mode=1, seed=3, dt_grid=run.time_interval / 300, n_path=16,
                          step_limit=64, stop_at_grid_end=False).trajectory(occ[1], traj_id=1, cap_steps=64, want_events=True)
        assert np.array_equal(res['events'][1, :64], one['events'][:64])
        done.append(f'{kernel} featured (field, events, times)')
        system.close()

    # ---- gather kernels: dense array (shipped example, replayed draws) and unit rows with the stencil disabled
    ex4, z = H.load_ref_case('hematite_4e')
    run4 = H.run_parameters(ex4)
    rng = H.rng_from_state_bytes(z['rnd_state_0'])
    occ4 = run4.initial_occupancy_from(rng)
    draws = H.draw_stream(rng, 200)
    system = K.KmcSystem(ctx, run4, ex4.P)
    ens = K.KmcEnsemble(system, np.array([occ4]), rng_mode=nat.RNG_REPLAY)
    res = ens.advance(200, draws=draws[None, :], want_events=True)
    ens.close()
    system.close()
    ref = O.KmcOracle(run4, ex4.P).trajectory(occ4, draws, want_events=True)
    n = min(int(res['steps_done'][0]), ref['n_steps'])
    assert np.array_equal(res['events'][0, :n], ref['events'][:n])
    done.append('kmc_step_kernel (dense array, replayed MT19937 draws)')

    # ---- BVO: 8 neighbour slots (stencil <1,1,8>)
    exb = H.load_example('bvo')
    scb = Supercell(exb.lattice, [3, 3, 1], [1, 1, 1])
    simb = exb.sim
    runb = K.RunParameters(exb.lattice, scb, scb.hop_neighbor_tables(), simb['temp'], 'full', 'full', simb['t_final'],
                           simb['time_interval'], [6, 0], {}, simb['relative_energies'], simb['external_field'])
    epb = EW.EwaldParameters(scb, exb.cfg['alpha'], exb.cfg['r_cut'], exb.cfg['k_cut'])
    pub, _ = EW.ewald_rows(ctx, epb, np.ascontiguousarray(scb.coordinates), 0, scb.n_per_cell)
    denseb = EW.ewald_expand(ctx, scb, pub, 0, scb.num_system_elements)
    occb = K.philox_initial_occupancy(runb.tables, 2, 6, seed=9)
    system = K.KmcSystem(ctx, runb, pub, layout='unit_rows')
    kw = dict(dt_grid=runb.time_interval / 50, n_path=8, step_limit=64, stop_at_grid_end=False)
    ens = K.KmcEnsemble(system, occb, rng_mode=nat.RNG_PHILOX, seed=9, refresh_interval=8, **kw)
    ens.advance_resident(64)
    got = ens.read()
    kern = ens.last_kernel()
    ens.close()
    system.close()
    ref = O.KmcOracle(runb, denseb, rng_mode=1, seed=9, **kw).ensemble(occb)
    assert np.array_equal(got['occupancy'], ref['occupancy'])
    done.append(f'{kern} (BVO, 8 slots)')

    # ---- MSD
    uw = np.stack([z['unwrapped_0'], z['unwrapped_1']])
    s4 = ex4.sim
    mp = M.MsdParameters(s4['n_dim'], s4['species_count'], 2, s4['t_final'], s4['time_interval'], s4['msd_t_final'],
                         s4['trim_length'], s4['temp'], s4['repr_time'], s4['repr_dist'])
    avg = M.species_avg_sd(ctx, uw[:, :400], 2, 400, mp.total_species, 60, mp.dist_conversion, mp.type_offsets)
    pos = uw[:, :400].reshape(2, 400, mp.total_species, 3) * mp.dist_conversion
    assert np.allclose(avg[:, :, 0], O.msd_sd(pos, 60).mean(axis=2), rtol=1e-11, atol=1e-9)
    done.append('msd kernels')
    for d in done:
        print('ok:', d)
    print(f'sanitize_run: {len(done)} kernel groups, {ctx.launch_count()} launches')


if __name__ == '__main__':
    main()
