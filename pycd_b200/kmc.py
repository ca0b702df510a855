"""Host side of the KMC run: run parameters, initial state, the ensemble driver.

Mirrors `Run.__init__` (PyCD/core.py:1685-1933, the parts the hot path needs),
`generate_initial_occupancy` (:2478-2528), `base_charge_config` (:2530-2538),
`preproduction` (:2566-2580) and the bookkeeping of `do_kmc_steps` (:2647-2917).  The
step loop itself (rates, selection, time advance, recording) runs in csrc/kmc.cu.
"""
import ctypes as C
import pickle
import random as _random

import numpy as np

from . import _native as nat
from . import constants
from .doping import DopingHooks
from .tables import HopTables


class RunParameters:
    """Everything static about a run of one carrier type (single-type only: the
    reference's mixed electron+hole bookkeeping is broken)."""

    def __init__(self, lattice, supercell, hop_neighbor_list, temp, ion_charge_type,
                 species_charge_type, t_final, time_interval, species_count,
                 initial_occupancy=None, relative_energies=None, external_field=None, doping=None):
        lat = lattice
        self.lattice, self.supercell = lattice, supercell
        self.species_count = np.asarray(species_count, dtype=int)
        active = [i for i, v in enumerate(self.species_count) if v > 0]
        if len(active) != 1:
            raise NotImplementedError(
                f'species_count={list(self.species_count)}: exactly one carrier type must be '
                'present (mixed electron+hole runs are broken in the reference itself, SURVEY F7)')
        self.species_index = active[0]
        self.species_type = lat.species_types[self.species_index]
        self.n_carriers = int(self.species_count[self.species_index])
        self.tables = HopTables(lat, supercell, hop_neighbor_list, self.species_type)
        self.n_proc = self.n_carriers * self.tables.nn  # core.py:1739
        self.ion_charge_type = ion_charge_type
        self.species_charge_type = species_charge_type
        self.q_carrier = float(lat.species_charge_list[species_charge_type][self.species_index])
        self.kT = temp * constants.K2AUTEMP  # core.py:1705
        self.vn = lat.vn
        self.t_final = t_final * constants.SEC2AUTIME  # core.py:1709-1710
        self.time_interval = time_interval * constants.SEC2AUTIME
        self.n_path = int(self.t_final / self.time_interval) + 1  # core.py:2657
        self.initial_occupancy = initial_occupancy or {}
        # relative site energies, core.py:1717-1730
        unit = np.zeros(lat.total_elements_per_unit_cell)
        head = 0
        for ei, name in enumerate(lat.element_types):
            n = int(lat.n_elements_per_unit_cell[ei])
            if lat.num_classes[ei] != 1:
                table = (relative_energies or {}).get('class_index', {}).get(name)
                if table is None:
                    raise KeyError(f"relative_energies['class_index']['{name}'] is required "
                                   f'({name} has {lat.num_classes[ei]} site classes)')
                unit[head:head + n] += [table[c] * constants.EV2HARTREE
                                        for c in lat.unit_cell_class_list[head:head + n]]
            head += n
        self.e_rel = np.tile(unit, supercell.num_cells)
        # lattice charges, core.py:2530-2538
        unit_q = np.array([lat.charge_types[ion_charge_type][lat.element_types[i]]
                           for i in lat.element_type_index_list], dtype=float)
        self.q_lat = np.tile(unit_q, supercell.num_cells)
        # doping hooks, core.py:1763-1781 (inactive unless num_dopants has a non-zero entry)
        self.doping = DopingHooks(lat, doping, relative_energies, ion_charge_type)
        # electric field, core.py:1749-1761
        self.field_active = 0
        self.field = np.zeros(3)
        self.field_mag = 0.0
        if external_field is not None:
            ef = external_field['electric']
            self.field_mag = ef['mag']
            if ef['active']:
                direction = np.asarray(ef['dir'], dtype=float)
                if ef['ld'] == 1:
                    direction = np.dot(direction, lat.lattice_matrix)
                self.field = ef['mag'] * (direction / np.linalg.norm(direction))
                self.field_active = 1

    def initial_energy(self, P, occupancy, alpha_per_angstrom, q_lat=None):
        """current_state_energy before the first step (core.py:2663-2667, 2773-2780):
        ewald_neut + sum_ij q_i q_j P_ij with q = lattice charges (q_lat: a doped trajectory's)
        + carriers.  alpha is the value material_run re-reads from precomputed_array.log
        (1/angstrom)."""
        sc = self.supercell
        alpha = alpha_per_angstrom / constants.ANG2BOHR  # core.py:768-769
        system_charge = float(np.dot(self.species_count,
                                     self.lattice.species_charge_list[self.species_charge_type]))
        ewald_neut = -(np.pi * system_charge ** 2 / (2 * sc.system_volume * alpha))
        q = np.array(self.q_lat if q_lat is None else q_lat, dtype=np.float64, copy=True)
        # core.py:2561-2564: `charge_list[sites] += q_c` is a fancy-index assignment, so a site that holds
        # several carriers (no site exclusion) is counted ONCE; replicated literally
        q[np.asarray(occupancy, dtype=int)] += self.q_carrier
        return ewald_neut + float(q @ (P @ q))

    # -- initial state -----------------------------------------------------------------
    def initial_occupancy_from(self, rng, dopant_site_indices=None):
        """generate_initial_occupancy (core.py:2483-2528): carriers started on dopant sites
        (site_charge_initiation, doped runs only), then explicit sites, then rng.sample over the
        element's sites in ascending index."""
        occ = []
        n = self.n_carriers
        if self.doping.active:
            occ, n = self.doping.initiation_sites(rng, self.species_type, n, dopant_site_indices or {})
        explicit = self.initial_occupancy.get(self.species_type) or []
        occ.extend(int(i) for i in explicit)
        n -= len(explicit)
        if n:
            pool = [int(s) for s in self.tables.sites if int(s) not in occ]
            occ.extend(rng.sample(pool, n))
        return occ


def write_initial_rnd_states(dst_path, n_traj, random_seed, only=None):
    """Run.preproduction, core.py:2573-2580: per-trajectory MT19937 states.  `only`: the trajectory
    indices this process writes (a rank's block under torchrun); the seeds are those of the full run."""
    rng = _random.Random()
    rng.seed(random_seed)
    seeds = [rng.random() for _ in range(n_traj)]
    for i in (range(n_traj) if only is None else only):
        d = dst_path / f'traj{i + 1}'
        d.mkdir(parents=True, exist_ok=True)
        rng.seed(seeds[i])
        with open(d / 'initial_rnd_state.dump', 'wb') as fh:
            pickle.dump(rng.getstate(), fh)


class _PlainUnpickler(pickle.Unpickler):
    """initial_rnd_state.dump holds random.getstate(): a tuple of ints / a tuple of ints / None -- no
    class may be instantiated while loading it."""

    def find_class(self, module, name):
        raise pickle.UnpicklingError(f'initial_rnd_state.dump must not reference {module}.{name}')


def load_rnd_state(path):
    rng = _random.Random()
    with open(path, 'rb') as fh:
        state = _PlainUnpickler(fh).load()
    if not (isinstance(state, tuple) and len(state) == 3 and isinstance(state[1], tuple)
            and all(isinstance(v, int) for v in state[1])):
        raise ValueError(f'{path}: not a random.getstate() dump')
    rng.setstate(state)
    return rng


def philox_initial_occupancy(tables, n_traj, n_carriers, seed, traj_id0=0):
    """Ensemble mode: C distinct carrier sites per trajectory, uniform without replacement,
    keyed by the GLOBAL trajectory id so that sharding over GPUs is invisible."""
    out = np.empty((n_traj, n_carriers), dtype=np.int32)
    for i in range(n_traj):
        g = np.random.Generator(np.random.Philox(key=[int(seed), int(traj_id0 + i)]))
        out[i] = tables.sites[g.choice(tables.n_centres, size=n_carriers, replace=False)]
    return out


class KmcSystem:
    """Device-resident static tables (pycd_kmc_system).  P may be a numpy array (copied
    to the device) or a torch CUDA tensor / device address (used in place).

    layout='dense': P is the full (N, N) array.  layout='unit_rows': P is (n_per_cell, N),
    the rows of unit cell 0 (ewald_rows(0, n_per_cell)); every element is reached through
    lattice translation, which needs pbc = [1, 1, 1]; the 7.2 MB table of the 10x10x10
    Hematite cell stays L2-resident instead of gathering from a 7.2 GB array."""

    def __init__(self, ctx, run, P, kT=None, field=None, layout='dense'):
        t = run.tables
        sc = run.supercell
        if layout not in ('dense', 'unit_rows'):
            raise ValueError("layout must be 'dense' or 'unit_rows'")
        if layout == 'unit_rows' and not bool(np.all(sc.pbc == 1)):
            raise ValueError("layout='unit_rows' needs pbc = [1, 1, 1]")
        self.layout = layout
        self.ctx, self.run = ctx, run
        self._keep = dict(
            site_centre=np.ascontiguousarray(t.site_centre, dtype=np.int32),
            site_class=np.ascontiguousarray(t.site_class, dtype=np.int32),
            neigh=np.ascontiguousarray(t.neigh, dtype=np.int32),
            hopvec=np.ascontiguousarray(t.hopvec, dtype=np.float64),
            lam=np.ascontiguousarray(t.lam, dtype=np.float64),
            vab=np.ascontiguousarray(t.vab, dtype=np.float64),
            e_rel=np.ascontiguousarray(run.e_rel, dtype=np.float64),
            q_lat=np.ascontiguousarray(run.q_lat, dtype=np.float64))
        if isinstance(P, np.ndarray):
            P = np.ascontiguousarray(P, dtype=np.float64)
            n = sc.num_system_elements
            want = (n, n) if layout == 'dense' else (sc.n_per_cell, n)
            if P.shape != want:
                raise ValueError(f'precomputed array has shape {P.shape}, expected {want}')
        self._P = P
        d = nat.KmcSystemDesc()
        d.n_sites = run.supercell.num_system_elements
        d.P = nat.ptr(P)
        d.n_centres = t.n_centres
        d.n_class = t.n_class
        d.nn = t.nn
        for k, v in self._keep.items():
            setattr(d, k, nat.ptr(v))
        d.q_carrier = run.q_carrier
        d.kT = float(run.kT if kT is None else kT)
        d.vn = float(run.vn)
        f = run.field if field is None else np.asarray(field, dtype=float)
        d.field[:] = [float(x) for x in f]
        d.field_active = int(run.field_active if field is None else bool(np.any(f != 0)))
        d.p_layout = nat.P_UNIT_ROWS if layout == 'unit_rows' else nat.P_DENSE
        d.n_basis = int(sc.n_per_cell)
        d.size[:] = [int(v) for v in sc.system_size]
        self._h = C.c_void_p()
        nat.check(nat.lib().pycd_kmc_system_create(ctx.handle, C.byref(d), C.byref(self._h)))

    @property
    def handle(self):
        if not self._h:
            raise nat.NativeError('KMC system already destroyed')
        return self._h

    def stencil_info(self):
        """(available, table bytes, reason when unavailable) of the lattice-stencil tables."""
        ok, nbytes, why = C.c_int32(), C.c_int64(), C.create_string_buffer(256)
        nat.check(nat.lib().pycd_kmc_system_stencil(self.handle, C.byref(ok), C.byref(nbytes), why, 256))
        return bool(ok.value), int(nbytes.value), why.value.decode()

    def v_lat(self):
        sc = self.run.supercell
        out = np.empty(sc.num_system_elements if self.layout == 'dense' else sc.n_per_cell)
        nat.check(nat.lib().pycd_kmc_system_vlat(self.handle, nat.ptr(out)))
        return out

    def close(self):
        if self._h:
            nat.lib().pycd_kmc_system_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KmcEnsemble:
    """A batch of independent trajectories on one GPU (pycd_kmc_ensemble)."""

    def __init__(self, system, occupancy0, dt_grid=None, n_path=None, step_limit=0,
                 stop_at_grid_end=True, rng_mode=nat.RNG_REPLAY, seed=0, traj_id0=0,
                 refresh_interval=1, kT_traj=None, field_traj=None, record_unwrapped=True, energy0=None,
                 doping=None, dt_grid_traj=None):
        """doping: None or one doping.TrajectoryDoping per trajectory (core.py:2723-2776): the
        trajectory's shifted site energies and the charges its dopant sites carry."""
        run = system.run
        self.system = system
        occ = np.ascontiguousarray(occupancy0, dtype=np.int32)
        if occ.ndim == 1:
            occ = occ[None, :]
        self.n_traj, self.n_carriers = occ.shape
        self.n_proc = self.n_carriers * run.tables.nn
        self.n_path = int(run.n_path if n_path is None else n_path)
        self.rng_mode = rng_mode
        d = nat.KmcEnsembleDesc()
        d.n_traj = self.n_traj
        d.traj_id0 = int(traj_id0)
        d.n_carriers = self.n_carriers
        d.occupancy0 = nat.ptr(occ)
        d.dt_grid = float(run.time_interval if dt_grid is None else dt_grid)
        d.n_path = self.n_path
        d.step_limit = int(step_limit)
        d.stop_at_grid_end = int(bool(stop_at_grid_end))
        d.rng_mode = int(rng_mode)
        d.seed = int(seed)
        d.refresh_interval = int(refresh_interval)
        keep = [occ]
        if kT_traj is not None:
            kT_traj = np.ascontiguousarray(kT_traj, dtype=np.float64)
            assert kT_traj.shape == (self.n_traj,)
            d.kT_traj = nat.ptr(kT_traj)
            keep.append(kT_traj)
        if field_traj is not None:
            field_traj = np.ascontiguousarray(field_traj, dtype=np.float64)
            assert field_traj.shape == (self.n_traj, 3)
            d.field_traj = nat.ptr(field_traj)
            keep.append(field_traj)
        if dt_grid_traj is not None:   # one time grid per trajectory (sweeps: one per condition)
            dt_grid_traj = np.ascontiguousarray(dt_grid_traj, dtype=np.float64)
            assert dt_grid_traj.shape == (self.n_traj,) and np.all(dt_grid_traj > 0)
            d.dt_grid_traj = nat.ptr(dt_grid_traj)
            keep.append(dt_grid_traj)
        self.has_energy = energy0 is not None
        if energy0 is not None:
            energy0 = np.ascontiguousarray(energy0, dtype=np.float64)
            assert energy0.shape == (self.n_traj,)
            d.energy0 = nat.ptr(energy0)
            keep.append(energy0)
        if doping is not None:
            if len(doping) != self.n_traj:
                raise ValueError('doping needs one entry per trajectory')
            n_sites = run.supercell.num_system_elements
            e_rel_traj = np.ascontiguousarray([t.e_rel for t in doping], dtype=np.float64)
            assert e_rel_traj.shape == (self.n_traj, n_sites)
            d.e_rel_traj = nat.ptr(e_rel_traj)
            keep.append(e_rel_traj)
            n_dop = max(len(t.sites) for t in doping)
            if n_dop:
                dsite = np.full((self.n_traj, n_dop), -1, dtype=np.int32)
                ddq = np.zeros((self.n_traj, n_dop))
                for i, t in enumerate(doping):
                    dsite[i, :len(t.sites)] = t.sites
                    ddq[i, :len(t.sites)] = t.dq
                d.n_dopant_max = n_dop
                d.dopant_site = nat.ptr(dsite)
                d.dopant_dq = nat.ptr(ddq)
                keep += [dsite, ddq]
        d.record_unwrapped = int(bool(record_unwrapped))
        self.record_unwrapped = bool(record_unwrapped)
        self._h = C.c_void_p()
        nat.check(nat.lib().pycd_kmc_ensemble_create(system.handle, C.byref(d), C.byref(self._h)))

    @property
    def handle(self):
        if not self._h:
            raise nat.NativeError('KMC ensemble already destroyed')
        return self._h

    def reset(self, occupancy0, traj_id0=0):
        """Re-arm with new initial sites (t = 0, empty grid) without re-allocating."""
        occ = np.ascontiguousarray(occupancy0, dtype=np.int32)
        if occ.shape != (self.n_traj, self.n_carriers):
            raise ValueError(f'occupancy must have shape {(self.n_traj, self.n_carriers)}')
        nat.check(nat.lib().pycd_kmc_ensemble_reset(self.handle, nat.ptr(occ), int(traj_id0)))

    def advance(self, max_steps, draws=None, want_events=False, want_times=False):
        """At most max_steps KMC steps per unfinished trajectory.  Returns a dict with
        n_active, steps_done[n_traj] and optionally events / times [n_traj, max_steps]."""
        max_steps = int(max_steps)
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            if draws.shape != (self.n_traj, 2 * max_steps):
                raise ValueError(f'draws must have shape {(self.n_traj, 2 * max_steps)}')
        events = np.full((self.n_traj, max_steps), -1, dtype=np.int32) if want_events else None
        times = np.zeros((self.n_traj, max_steps)) if want_times else None
        steps_done = np.zeros(self.n_traj, dtype=np.int64)
        n_active = C.c_int64()
        nat.check(nat.lib().pycd_kmc_advance(self.handle, max_steps, nat.ptr(draws), nat.ptr(events),
                                             nat.ptr(times), nat.ptr(steps_done), C.byref(n_active)))
        return {'n_active': n_active.value, 'steps_done': steps_done, 'events': events, 'times': times}

    def advance_resident(self, max_steps):
        """Philox mode without any per-call host buffers (bench inner loop)."""
        n_active = C.c_int64()
        nat.check(nat.lib().pycd_kmc_advance(self.handle, int(max_steps), None, None, None, None,
                                             C.byref(n_active)))
        return n_active.value

    def advance_async(self, max_steps):
        """Enqueue `max_steps` KMC steps without waiting (Philox mode, no per-step outputs); see wait()."""
        nat.check(nat.lib().pycd_kmc_advance_async(self.handle, int(max_steps)))

    def wait(self):
        """Block until the launches of advance_async have finished; number of unfinished trajectories."""
        n_active = C.c_int64()
        nat.check(nat.lib().pycd_kmc_wait(self.handle, C.byref(n_active)))
        return n_active.value

    def read(self, unwrapped=True, rates=False, out=None):
        """State read-back.  `out` may carry pre-allocated (e.g. pinned) arrays from a previous
        call (the returned dict) to avoid re-allocating the result buffers."""
        nt, C_ = self.n_traj, self.n_carriers
        if out is None:
            out = {'n_steps': np.zeros(nt, dtype=np.int64), 'time': np.zeros(nt),
                   'occupancy': np.zeros((nt, C_), dtype=np.int32), 'drift': np.zeros((nt, C_, 3)),
                   'near_tie': np.zeros(nt, dtype=np.int64), 'clamped': np.zeros(nt, dtype=np.int64),
                   'unwrapped': None, 'rates': None}
        uw = None
        if unwrapped and self.record_unwrapped:
            uw = out.get('unwrapped')
            if uw is None:
                uw = np.empty((nt, self.n_path, 3 * C_))
        rt = None
        if rates:
            rt = out.get('rates')
            if rt is None:
                rt = np.empty((nt, self.n_proc))
        nat.check(nat.lib().pycd_kmc_read(self.handle, nat.ptr(uw), nat.ptr(out['n_steps']),
                                          nat.ptr(out['time']), nat.ptr(out['occupancy']),
                                          nat.ptr(out['drift']), nat.ptr(out['near_tie']),
                                          nat.ptr(out['clamped']), nat.ptr(rt)))
        out['unwrapped'] = uw
        out['rates'] = rt
        return out

    def read_begin(self, out):
        """Pipelined read-back into the (pinned) buffers of `out` (pinned_buffers()): returns at once;
        reset() + advance of the next batch overlap the copy; read_end() completes it."""
        uw = out.get('unwrapped') if self.record_unwrapped else None
        nat.check(nat.lib().pycd_kmc_read_begin(self.handle, nat.ptr(uw), nat.ptr(out['n_steps']),
                                                nat.ptr(out['time']), nat.ptr(out['occupancy']),
                                                nat.ptr(out['drift']), nat.ptr(out['near_tie']),
                                                nat.ptr(out['clamped'])))
        return out

    def read_end(self):
        nat.check(nat.lib().pycd_kmc_read_end(self.handle))

    def pinned_buffers(self):
        """Result buffers in page-locked host memory, to be passed as read(out=...)."""
        nt, C_ = self.n_traj, self.n_carriers
        return {'n_steps': nat.pinned_empty(nt, np.int64), 'time': nat.pinned_empty(nt),
                'occupancy': nat.pinned_empty((nt, C_), np.int32), 'drift': nat.pinned_empty((nt, C_, 3)),
                'near_tie': nat.pinned_empty(nt, np.int64), 'clamped': nat.pinned_empty(nt, np.int64),
                'unwrapped': nat.pinned_empty((nt, self.n_path, 3 * C_)) if self.record_unwrapped else None,
                'rates': None}

    def read_energy(self):
        """(energy_traj, delG0_traj), each (n_traj, n_path) (output_data energy / delg_0)."""
        e = np.empty((self.n_traj, self.n_path))
        g = np.empty((self.n_traj, self.n_path))
        nat.check(nat.lib().pycd_kmc_read_energy(self.handle, nat.ptr(e), nat.ptr(g)))
        return e, g

    def last_kernel(self):
        """Name of the step kernel the last advance launched."""
        buf = C.create_string_buffer(128)
        nat.check(nat.lib().pycd_kmc_last_kernel(self.handle, buf, 128))
        return buf.value.decode()

    def unwrapped_device_ptr(self):
        p = C.c_void_p()
        nat.check(nat.lib().pycd_kmc_unwrapped_device(self.handle, C.byref(p)))
        return p.value

    def close(self):
        if self._h:
            nat.lib().pycd_kmc_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_replay(system, rngs, occupancy0, chunk_steps=32768, want_times=True, want_events=False,
               max_total_steps=None, energy0=None, doping=None):
    """Runs trajectories to the end of their time grid, feeding each one the continuing
    stream of its own Python MT19937 generator (u1 then u2 per step, core.py:2799-2802).
    Returns (state dict from KmcEnsemble.read, list of per-trajectory time arrays incl. the
    leading 0.0, list of per-trajectory event arrays or None)."""
    ens = KmcEnsemble(system, occupancy0, rng_mode=nat.RNG_REPLAY, refresh_interval=1, energy0=energy0,
                      doping=doping)
    n_traj = ens.n_traj
    times = [[np.zeros(1)] for _ in range(n_traj)]
    events = [[] for _ in range(n_traj)]
    active = np.ones(n_traj, dtype=bool)
    total = 0
    while active.any():
        draws = np.zeros((n_traj, 2 * chunk_steps))
        for i in range(n_traj):
            if active[i]:
                rnd = rngs[i].random
                draws[i] = [rnd() for _ in range(2 * chunk_steps)]
        res = ens.advance(chunk_steps, draws=draws, want_events=want_events, want_times=want_times)
        for i in range(n_traj):
            n = int(res['steps_done'][i])
            if want_times and n:
                times[i].append(res['times'][i, :n].copy())
            if want_events and n:
                events[i].append(res['events'][i, :n].copy())
            if n < chunk_steps:
                active[i] = False
        total += chunk_steps
        if res['n_active'] == 0 or (max_total_steps and total >= max_total_steps):
            break
    state = ens.read()
    state['last_kernel'] = ens.last_kernel()
    if energy0 is not None:
        state['energy_grid'], state['dg0_grid'] = ens.read_energy()
    ens.close()
    t_out = [np.concatenate(t) for t in times] if want_times else None
    e_out = [np.concatenate(e) if e else np.zeros(0, dtype=np.int32) for e in events] \
        if want_events else None
    return state, t_out, e_out


def drift_mobility(drift, field, field_mag):
    """compute_drift_mobility, core.py:2054-2059: (n_traj, C) in cm2/V.s."""
    au = np.dot(drift, field) / field_mag ** 2
    return au * (constants.BOHR2CM ** 2 * constants.SEC2AUTIME * constants.V2AUPOT)
