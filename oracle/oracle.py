"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/libpycd_oracle.so (pycd_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  pycd_b200/ never imports this.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / 'libpycd_oracle.so'


def build(force=False):
    src = HERE / 'pycd_oracle.c'
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        env = dict(os.environ)
        env.pop('CC', None)
        res = subprocess.run(['make', '-C', str(HERE)] + (['-B'] if force else []),
                             capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError('oracle build failed:\n' + res.stdout + res.stderr)
    return LIB


class KmcParams(C.Structure):
    _fields_ = [('n_sites', C.c_int64), ('P', C.c_void_p), ('site_centre', C.c_void_p),
                ('site_class', C.c_void_p), ('nn', C.c_int32), ('neigh', C.c_void_p),
                ('hopvec', C.c_void_p), ('lam', C.c_void_p), ('vab', C.c_void_p),
                ('e_rel', C.c_void_p), ('q_lat', C.c_void_p), ('v_lat', C.c_void_p),
                ('q_carrier', C.c_double), ('kT', C.c_double), ('vn', C.c_double),
                ('field', C.c_double * 3), ('field_active', C.c_int32),
                ('n_carriers', C.c_int32), ('dt_grid', C.c_double), ('n_path', C.c_int64),
                ('step_limit', C.c_int64), ('stop_at_grid_end', C.c_int32),
                ('literal', C.c_int32), ('rng_mode', C.c_int32), ('seed', C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        h = C.CDLL(str(LIB))
        h.oracle_min_image.restype = None
        h.oracle_pairwise.restype = None
        h.oracle_ewald_literal.restype = C.c_int64
        h.oracle_ewald_literal.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        h.oracle_pairwise.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        h.oracle_philox_uniforms.restype = None
        h.oracle_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        h.oracle_kmc_trajectory.restype = C.c_int64
        h.oracle_kmc_trajectory.argtypes = [C.POINTER(KmcParams), C.c_uint64, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_double, C.c_void_p, C.c_void_p]
        h.oracle_kmc_ensemble.restype = C.c_int64
        h.oracle_kmc_ensemble.argtypes = [C.POINTER(KmcParams), C.c_int64, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        h.oracle_vlat.restype = None
        h.oracle_vlat.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        h.oracle_msd_sd.restype = None
        h.oracle_msd_sd.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
        _lib = h
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def pairwise(coords, cell, cellinv, pbc):
    """(N,N,3) minimum-image vectors, core.py:571-594 + 304-361."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    n = len(coords)
    out = np.empty((n, n, 3))
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    cellinv = np.ascontiguousarray(cellinv, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.int32)
    lib().oracle_pairwise(_p(coords), n, _p(cell), _p(cellinv), _p(pbc), _p(out))
    return out


def ewald_literal(pair, recip, volume, alpha, r_cut, k_cut, eps, k_max, parts=False):
    """core.py:799-878 + 1594-1602 + 1659-1661 on an (N,N,3) pair-vector array."""
    pair = np.ascontiguousarray(pair, dtype=np.float64)
    n = pair.shape[0]
    recip = np.ascontiguousarray(recip, dtype=np.float64)
    k_max = np.ascontiguousarray(k_max, dtype=np.int32)
    out = np.empty((n, n))
    re = np.empty((n, n)) if parts else None
    fo = np.empty((n, n)) if parts else None
    keff = lib().oracle_ewald_literal(_p(pair), n, _p(recip), float(volume), float(alpha), float(r_cut),
                                      float(k_cut), float(eps), _p(k_max), _p(out), _p(re), _p(fo))
    return (out, keff, re, fo) if parts else (out, keff)


def philox_uniforms(seed, traj, step):
    u = np.empty(2)
    lib().oracle_philox_uniforms(int(seed), int(traj), int(step), _p(u))
    return u


def vlat(P, q_lat):
    P = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(q_lat, dtype=np.float64)
    out = np.empty(len(q))
    lib().oracle_vlat(_p(P), _p(q), len(q), _p(out))
    return out


class KmcOracle:
    """Holds the flat tables (same arrays the CUDA path uploads) and runs trajectories on
    the CPU, literally (O(N) dot with the charge vector) or in the gather form."""

    def __init__(self, run, P, literal=False, kT=None, field=None, dt_grid=None, n_path=None,
                 step_limit=0, stop_at_grid_end=True, rng_mode=0, seed=0, e_rel=None, q_lat=None):
        # e_rel / q_lat: a doped trajectory's site energies and lattice charges (core.py:2723-2764,
        # 2553-2557) in place of the undoped ones
        t = run.tables
        self.run = run
        self.a = dict(
            P=np.ascontiguousarray(P, dtype=np.float64),
            site_centre=np.ascontiguousarray(t.site_centre, dtype=np.int32),
            site_class=np.ascontiguousarray(t.site_class, dtype=np.int32),
            neigh=np.ascontiguousarray(t.neigh, dtype=np.int32),
            hopvec=np.ascontiguousarray(t.hopvec, dtype=np.float64),
            lam=np.ascontiguousarray(t.lam, dtype=np.float64),
            vab=np.ascontiguousarray(t.vab, dtype=np.float64),
            e_rel=np.ascontiguousarray(run.e_rel if e_rel is None else e_rel, dtype=np.float64),
            q_lat=np.ascontiguousarray(run.q_lat if q_lat is None else q_lat, dtype=np.float64))
        self.a['v_lat'] = vlat(self.a['P'], self.a['q_lat'])
        p = KmcParams()
        p.n_sites = run.supercell.num_system_elements
        for k, v in self.a.items():
            setattr(p, k, _p(v))
        p.nn = t.nn
        p.q_carrier = run.q_carrier
        p.kT = float(run.kT if kT is None else kT)
        p.vn = float(run.vn)
        f = run.field if field is None else np.asarray(field, dtype=float)
        p.field[:] = [float(x) for x in f]
        p.field_active = int(run.field_active if field is None else bool(np.any(f != 0)))
        p.n_carriers = run.n_carriers
        p.dt_grid = float(run.time_interval if dt_grid is None else dt_grid)
        p.n_path = int(run.n_path if n_path is None else n_path)
        p.step_limit = int(step_limit)
        p.stop_at_grid_end = int(bool(stop_at_grid_end))
        p.literal = int(bool(literal))
        p.rng_mode = int(rng_mode)
        p.seed = int(seed)
        self.p = p
        self.n_proc = run.n_carriers * t.nn

    def trajectory(self, occ, draws=None, traj_id=0, cap_steps=0, want_times=False, want_events=False,
                   energy0=None):
        C_ = self.p.n_carriers
        occ = np.ascontiguousarray(occ, dtype=np.int32).copy()
        assert occ.shape == (C_,)
        n_draw_steps = 0
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            n_draw_steps = len(draws) // 2
            cap_steps = cap_steps or n_draw_steps
        uw = np.empty((self.p.n_path, 3 * C_))
        times = np.zeros(cap_steps + 1) if want_times else None
        events = np.full(cap_steps, -1, dtype=np.int32) if want_events else None
        drift = np.zeros((C_, 3))
        rates0 = np.zeros(self.n_proc)
        dg0 = np.zeros(self.n_proc)
        clamped = C.c_int64(0)
        energy = C.c_double(0.0)
        e_grid = np.zeros(self.p.n_path) if energy0 is not None else None
        g_grid = np.zeros(self.p.n_path) if energy0 is not None else None
        n = lib().oracle_kmc_trajectory(C.byref(self.p), int(traj_id), _p(occ), _p(draws), n_draw_steps,
                                        _p(uw), _p(times), _p(events), int(cap_steps), _p(drift),
                                        _p(rates0), C.byref(clamped), C.byref(energy), _p(dg0),
                                        float(energy0 or 0.0), _p(e_grid), _p(g_grid))
        return {'n_steps': int(n), 'occupancy': occ, 'unwrapped': uw,
                'times': times[:n + 1] if want_times else None,
                'events': events[:n] if want_events else None, 'drift': drift, 'rates0': rates0,
                'dg0_first': dg0, 'clamped': clamped.value, 'energy_change': energy.value,
                'energy_grid': e_grid, 'dg0_grid': g_grid}

    def ensemble(self, occ, traj_id0=0, draws=None, want_unwrapped=True, n_threads=0):
        C_ = self.p.n_carriers
        occ = np.ascontiguousarray(occ, dtype=np.int32).copy()
        n_traj = occ.shape[0]
        n_draw_steps = 0
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            n_draw_steps = draws.shape[1] // 2
        uw = np.empty((n_traj, self.p.n_path, 3 * C_)) if want_unwrapped else None
        nsteps = np.zeros(n_traj, dtype=np.int64)
        drift = np.zeros((n_traj, C_, 3))
        total = lib().oracle_kmc_ensemble(C.byref(self.p), n_traj, int(traj_id0), _p(occ), _p(draws),
                                          n_draw_steps, _p(uw), _p(nsteps), _p(drift), int(n_threads))
        return {'total_steps': int(total), 'n_steps': nsteps, 'occupancy': occ, 'unwrapped': uw,
                'drift': drift}


def msd_sd(pos, n_msd):
    """pos: (n_traj, n_path, C, 3) scaled positions -> sd (n_traj, n_msd, C), core.py:2996-3007."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    n_traj, n_path, C_, _ = pos.shape
    out = np.empty((n_traj, n_msd, C_))
    lib().oracle_msd_sd(_p(pos), n_traj, n_path, C_, n_msd, _p(out))
    return out


def msd_analysis(unwrapped, species_count, n_msd, time_interval_au, time_conversion, dist_conversion,
                 trim, temp_K, n_dim, linregress=None):
    """Analysis.compute_msd, core.py:2990-3071 on unwrapped (n_traj, n_path, 3C) [bohr].
    Returns msd_data (n_msd, 1+n_types), sem_data, slopes, diffusivity list, sem list."""
    from pycd_b200 import constants  # values only
    unwrapped = np.asarray(unwrapped)
    n_traj, n_path, c3 = unwrapped.shape
    C_ = c3 // 3
    pos = unwrapped.reshape(n_traj, n_path, C_, 3) * dist_conversion
    sd = msd_sd(pos, n_msd)
    counts = [int(c) for c in species_count if c != 0]
    avg = np.zeros((n_traj, n_msd, len(counts)))
    start = 0
    for k, c in enumerate(counts):
        avg[:, :, k] = np.mean(sd[:, :, start:start + c], axis=2)
        start += c
    msd = np.zeros((n_msd, len(counts) + 1))
    msd[:, 0] = np.arange(n_msd) * time_interval_au * time_conversion
    msd[:, 1:] = np.mean(avg, axis=0)
    sem = np.std(avg, axis=0) / np.sqrt(n_traj)
    slopes = np.zeros((n_traj, len(counts)))
    x = msd[trim:-trim, 0]
    for k in range(len(counts)):
        for tr in range(n_traj):
            y = avg[tr, trim:-trim, k]
            if linregress is not None:
                slopes[tr, k] = linregress(x, y)[0]
            else:
                xm, ym = x.mean(), y.mean()
                slopes[tr, k] = np.dot(x - xm, y - ym) / np.dot(x - xm, x - xm)
    kBT = constants.KB * temp_K / constants.EV2J
    factor = constants.ANG2CM ** 2 * constants.SEC2NS / (2 * n_dim) / kBT
    diff = [float(np.mean(slopes[:, k]) * factor) for k in range(len(counts))]
    diff_sem = [float(np.std(slopes[:, k]) / np.sqrt(n_traj) * factor) for k in range(len(counts))]
    return {'msd_data': msd, 'sem_data': sem, 'species_avg_sd': avg, 'slopes': slopes,
            'diffusivity': diff, 'diffusivity_sem': diff_sem}
