#!/usr/bin/env python
"""bench.py -- KMC steps/s of the B200 hot path on the BASELINE.json workload.

Workload (config.workload): "Hematite 10x10x10 supercell, 64 carriers, 4096-trajectory
ensemble sharded over 8xB200" = 512 trajectories per GPU (weak scaling: per-GPU work is
fixed, trajectories are independent, no data-path collective).  One bench "step" = one
launch of the step kernel advancing every trajectory of the rank by --kmc-steps KMC steps
(rate evaluation + selection + time advance + recording).  Synthetic data: POSCAR geometry
of the reference's Hematite example, Ewald array built on the GPU (the reference cannot
build it at this size), Philox4x32-10 draws, random distinct initial sites.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / 'oracle', ROOT / 'tests'):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLD = ROOT / 'tests' / 'golden'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, nargs=3, default=[10, 10, 10])
    ap.add_argument('--carriers', type=int, default=64)
    ap.add_argument('--traj-per-gpu', type=int, default=512)
    ap.add_argument('--kmc-steps', type=int, default=16384, help='KMC steps per trajectory per bench step')
    ap.add_argument('--refresh', type=int, default=256,
                    help='1 = stateless rate evaluation; R>1 = incremental updates, full re-gather every R')
    ap.add_argument('--n-path', type=int, default=101, help='rows of the recorded time grid')
    ap.add_argument('--cpu-seconds', type=float, default=20.0, help='budget of the cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parity-budget', type=int, default=6_000_000,
                    help='KMC steps the CPU oracle may spend re-running trajectories of the timed ensemble')
    ap.add_argument('--skip-msd', action='store_true')
    ap.add_argument('--no-flush', action='store_true', help='diagnostic: skip the L2 flush between iterations')
    ap.add_argument('--no-clocks', action='store_true', help='diagnostic: do not sample nvidia-smi clocks')
    ap.add_argument('--dense', action='store_true', help='step kernel gathers from the dense N x N array')
    ap.add_argument('--config', type=int, default=3, choices=[2, 3, 4, 5],
                    help='BASELINE.json configuration: 3 = the headline line (with bounded legs of 2, 4 and 5 '
                         'attached as `configs`); 2 / 4 / 5 = that configuration alone at full size')
    ap.add_argument('--skip-configs', action='store_true', help='headline line without the cfg 2 / 4 / 5 legs')
    ap.add_argument('--sweep-traj', type=int, default=1024, help='--config 5: trajectories per condition (total)')
    ap.add_argument('--sweep-common-grid', action='store_true',
                    help='--config 5: one time grid for all conditions (hot conditions then take ~100x more steps)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
def build_problem(args):
    """Geometry + tables of the benchmark system (host side, seconds)."""
    import yaml
    from pycd_b200.ewald import EwaldParameters
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Lattice, Supercell
    d = GOLD / 'hematite'
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    sim = yaml.safe_load(open(d / 'simulation_parameters.yml'))
    lat = Lattice(SimpleNamespace(**cfg))
    sc = Supercell(lat, args.size, [1, 1, 1])
    hop = sc.hop_neighbor_tables()
    run = RunParameters(lat, sc, hop, 300, 'full', 'full', sim['t_final'], sim['time_interval'],
                        [args.carriers, 0], {}, sim['relative_energies'], sim['external_field'])
    ep = EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    return lat, sc, run, ep


def b_step_bytes(n_proc, n_carriers):
    """Algorithmic bytes of one KMC step of one trajectory in the stateless gather
    formulation: 8*n_proc*(2C+6) + 4*n_proc."""
    return 8 * n_proc * (2 * n_carriers + 6) + 4 * n_proc


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """SM clock / throttle reasons sampled every 20 ms DURING the timed region (NVML in a thread;
    falls back to one `nvidia-smi` query when the bindings are missing)."""
    BITS = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}

    def __init__(self, index, period=0.02):
        self.sm, self.mx, self.reasons, self.power = [], None, set(), []
        self.index, self.period = index, period
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def start(self):
        """Begin sampling (the NVML set-up above takes milliseconds and stays outside the timed region)."""
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(self.period)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self._stop.set()
        self.thread.join(timeout=2)
        try:
            self._sample()
        except Exception:
            pass
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.mx,
                'samples': len(self.sm), 'power_w_max': max(self.power) if self.power else None,
                'reasons': sorted(self.reasons), 'how': 'NVML every 20 ms inside the timed region'}

    def _smi_once(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                                  '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=20).stdout
            r = [x.strip() for x in out.strip().split(',')]
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            return {'sm_mhz': float(r[0]), 'sm_max_mhz': float(r[1]), 'samples': 1,
                    'reasons': [n for n, v in zip(names, r[2:]) if v.lower().startswith('active')],
                    'how': 'one nvidia-smi query after the timed region (NVML bindings unavailable)'}
        except Exception:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'samples': 0, 'reasons': ['nvidia-smi unavailable']}


def measured_peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        d = json.load(open(f))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def step_floor():
    """Denominators of the step kernel's roofline from the committed microbenchmark of this device class
    (tools/step_floor.cu -> profiles/r02_step_floor.json): L2 random 32-byte-sector gather peak and the
    latency floor of one KMC step's dependency chain in the two-warp stencil kernel."""
    f = ROOT / 'profiles' / 'r02_step_floor.json'
    d = json.load(open(f)) if f.exists() else {}
    g = lambda k, v: float(d.get(k, v))
    dfma, dadd, dmma = g('lat_dfma_cycles', 8.4), g('lat_dadd_cycles', 8.2), g('lat_dmma_m8n8k4_operand_chain_cycles', 26.8)
    lds, bar, redux = g('lat_lds_dependent_cycles', 34.0), g('lat_bar_sync_64_threads_cycles', 16.6), g('lat_redux_add_cycles', 44.2)
    gather = g('lat_gather_3x32_sectors_1_warp_per_sm_cycles', 535.4)
    terms = {
        'rates: 22 dependent FP64 operations (Marcus argument, library exp sequence, pow correction, prefactor)': 22 * dfma,
        'barrier + LDS + 3-level lane prefix': bar + lds + 3 * dadd,
        'warp scan: 2 dependent DMMA + 2 FP64 operations': 2 * dmma + 2 * dfma,
        'selection: threshold, compare, integer reduction (REDUX)': 2 * dfma + 30 + redux,
        'table lookups of the selected process (LDS) + address arithmetic': lds + 15,
        'three divergent 32-byte gathers per lane from the L2-resident table': gather,
        'sum over carriers of the moved carrier\'s new terms: FMA + 4 chained DMMA + add + DMMA': dfma + 5 * dmma + dadd,
        'exchange between the two warps: STS + barrier + LDS + 2 adds': bar + lds + 2 * dadd,
        'patch of the cached sums': 2 * dfma,
    }
    return {'l2_gather_peak_gb_per_s': g('l2_gather_peak_gb_per_s', 9270.1),
            'chain_cycles': float(sum(terms.values())), 'chain_terms': {k: round(v, 1) for k, v in terms.items()},
            'source': 'measured (tools/step_floor.cu on a B200 of this pool, profiles/r02_step_floor.json)'
                      if d else 'fallback constants (profiles/r02_step_floor.json missing)'}


def profiled_traffic():
    """dram bytes per launch of the step kernel from the committed ncu capture, if any."""
    f = ROOT / 'profiles' / 'kmc_step_traffic.json'
    if f.exists():
        try:
            return json.load(open(f)).get('dram_bytes_per_launch')
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------
def cpu_reference_rate(run, P_host, occ, args, n_traj, steps_per_traj, n_path, dt_grid, seed,
                       literal=True, threads=None):
    """The oracle (C port of the reference's step loop) on the host cores: one independent
    trajectory per thread at a time, literal = the reference's O(N) dot per process."""
    import oracle as O
    threads = threads or host_cores()
    orc = O.KmcOracle(run, P_host, literal=literal, dt_grid=dt_grid, n_path=n_path,
                      step_limit=steps_per_traj, stop_at_grid_end=False, rng_mode=1, seed=seed)
    t0 = time.perf_counter()
    res = orc.ensemble(occ[:n_traj], want_unwrapped=False, n_threads=threads)
    dt = time.perf_counter() - t0
    return res['total_steps'] / dt, dt, res['total_steps'], threads


def ewald_cpu_baseline(sc, ep, n_sample=32):
    """The literal Ewald evaluation of the reference (core.py:799-878: one cos per (pair, k) term, all
    K_eff half-space vectors) timed by the oracle on n_sample sites of the bench supercell with all host
    threads, and extrapolated to the N x N array (cost is proportional to N^2 K_eff)."""
    import oracle as O
    n = sc.num_system_elements
    sites = np.random.default_rng(7).choice(n, size=n_sample, replace=False)
    pair = O.pairwise(np.ascontiguousarray(sc.coordinates[sites]), sc.cell_matrix, sc.cell_matrix_inv, sc.pbc)
    t0 = time.perf_counter()
    _, keff = O.ewald_literal(pair, sc.reciprocal_lattice_matrix, sc.system_volume, ep.alpha, ep.r_cut,
                              ep.k_cut, ep.dielectric, ep.k_max)
    dt = time.perf_counter() - t0
    ns_term = dt / (n_sample * n_sample * keff) * 1e9
    return {'kind': 'port', 'cores': host_cores(),
            'sample': f'{n_sample} x {n_sample} site pairs x {keff} k-vectors, literal cos sum (oracle), {dt:.1f} s',
            'ns_per_pair_k_term': ns_term,
            'seconds_full_array_extrapolated': ns_term * 1e-9 * n * n * keff,
            'note': 'extrapolated in proportion to N^2 K_eff; the reference itself (numpy temporaries, one '
                    'core) measures 20.4 ns per term (SURVEY 8a)'}


def parity_check(run, P_host, occ, state, n_traj, steps, n_path, dt_grid, seed, traj_id0):
    """The oracle (gather form, all host threads) on the first n_traj trajectories of the timed ensemble,
    `steps` KMC steps each; compared with the state the GPU reached (read back after the timed region)."""
    import oracle as O
    orc = O.KmcOracle(run, P_host, literal=False, dt_grid=dt_grid, n_path=n_path, step_limit=steps,
                      stop_at_grid_end=False, rng_mode=1, seed=seed)
    t0 = time.perf_counter()
    ref = orc.ensemble(occ[:n_traj], traj_id0=traj_id0, want_unwrapped=True, n_threads=host_cores())
    dt = time.perf_counter() - t0
    same = [bool(np.array_equal(ref['occupancy'][i], state['occupancy'][i]) and
                 ref['n_steps'][i] == state['n_steps'][i] and
                 np.array_equal(ref['unwrapped'][i], state['unwrapped'][i])) for i in range(n_traj)]
    return {'checked': n_traj, 'equal': int(sum(same)), 'kmc_steps_per_trajectory': int(steps),
            'what': 'final carrier sites, step counts and displacement grids of the timed ensemble (warm-up + '
                    'timed launches) vs the CPU oracle (O(C) gather restatement of core.py:1989-2050, 2787-2861) '
                    'run from the same initial sites and Philox keys',
            'rate': ref['total_steps'] / dt, 'seconds': dt}


def size_cpu_sample(run, P_host, occ, args, n_path, dt_grid, seed, literal):
    """Pick (#trajectories, steps) so that the sample costs about args.cpu_seconds."""
    cores = host_cores()
    probe_steps = 4 if literal else 64
    rate1, dt, _, _ = cpu_reference_rate(run, P_host, occ, args, 1, probe_steps, n_path, dt_grid, seed,
                                         literal=literal, threads=1)
    n_traj = min(cores, len(occ))
    steps = int(max(probe_steps, min(20000, rate1 * args.cpu_seconds)))
    return n_traj, steps, rate1


# ---------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rc = 0
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference' and rank != 0:
        return 0

    import torch
    has_gpu = torch.cuda.is_available()
    dist = None
    if world > 1 and args.impl == 'ours':
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        # NCCL may print its version banner on stdout when the communicator is created;
        # stdout carries exactly one JSON line, so the banner is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    if not has_gpu:
        if args.impl == 'reference':
            return reference_arm_small(args)
        raise SystemExit('bench.py --impl ours needs a B200; there is no CPU fallback')

    from pycd_b200 import _native as nat
    from pycd_b200 import constants
    from pycd_b200 import ewald as EW
    from pycd_b200 import kmc as K
    from pycd_b200 import msd as M
    if nat.needs_build():
        nat.build()

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ctx = nat.default_context(local_rank)
    lat, sc, run, ep = build_problem(args)
    N = sc.num_system_elements
    C_, nn = run.n_carriers, run.tables.nn
    n_proc = C_ * nn
    nt = args.traj_per_gpu
    seed = 2

    # ---- Ewald precompute: each rank builds its row block (unit-cell rows + translation
    # expansion), NCCL all-gather into the full N x N array on every GPU ---------------
    P = torch.empty((N, N), dtype=torch.float64, device=dev)
    coords = torch.from_numpy(np.ascontiguousarray(sc.coordinates)).to(dev)
    r0, r1 = (rank * N) // world, ((rank + 1) * N) // world
    p_unit = torch.empty((sc.n_per_cell, N), dtype=torch.float64, device=dev)
    # warm-up on the shipped 2x2x1 cell: the CUDA driver loads a kernel's code at its first launch, and both
    # tile shapes + the expansion are used below
    from pycd_b200 import constants as _c
    from pycd_b200.lattice import Supercell as _Supercell
    _sc0 = _Supercell(lat, [2, 2, 1], [1, 1, 1])
    _ep0 = EW.EwaldParameters(_sc0, ep.alpha * _c.ANG2BOHR, ep.r_cut / _c.ANG2BOHR, ep.k_cut * _c.ANG2BOHR)
    _c0 = np.ascontiguousarray(_sc0.coordinates)
    _pu0, _ = EW.unit_cell_rows(ctx, _ep0, _c0, method='cells')
    EW.unit_cell_rows(ctx, _ep0, _c0, method='rows')
    EW.ewald_expand(ctx, _sc0, _pu0, 0, _sc0.num_system_elements)
    EW.ewald_rows(ctx, _ep0, _c0, 0, _sc0.num_system_elements)
    if dist:   # NCCL sets up its channels on the first collective of each kind and size class
        warm = torch.zeros_like(p_unit)
        dist.all_reduce(warm)
        del warm
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    # ---- first, for the record and to bring the idle GPU up to its clocks (the host spent seconds building tables):
    # the same rows through the DMMA kernel, k list split over the ranks + one all-reduce -- the FP64 roofline
    # kernel of the general path (partial periodic boundaries, dense rows)
    p_chk = torch.empty_like(p_unit)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t1 = time.perf_counter()
    _, est_rows = EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=p_chk.data_ptr(),
                                k_part=rank, k_parts=world)
    t_reduce = 0.0
    if dist:
        tr = time.perf_counter()
        dist.all_reduce(p_chk)
        torch.cuda.synchronize()
        t_reduce = time.perf_counter() - tr
    rows_seconds = time.perf_counter() - t1
    # ---- the timed precompute: rows of unit cell 0 by the class-factorised sum (csrc/ewald_cells.cu: one pass
    # over the k vectors for the n_per_cell^2 basis pairs + a Fourier transform over the cells), then the dense
    # N x N array by lattice translation.  Every rank does both itself: at 2 ms per 7.2 GB the local expansion
    # is cheaper than an NVLink all-gather of row blocks (measured 22 ms at 8 GPUs), so the symmetric path needs
    # NO collective; the row-sharded dense evaluation + all-gather is the cfg-4 leg (`configs.cfg4_ewald`).
    def precompute():
        t0 = time.perf_counter()
        _, st = EW.unit_cell_rows(ctx, ep, coords.data_ptr(), out=p_unit.data_ptr())
        ta = time.perf_counter()
        EW.ewald_expand(ctx, sc, p_unit.data_ptr(), 0, N, out=P.data_ptr())
        ms = ctx.last_kernel_ms(nat.KC_EWALD_EXPAND)
        torch.cuda.synchronize()
        tb = time.perf_counter()
        if os.environ.get('PYCD_BENCH_DEBUG'):
            print(f'precompute: unit rows {1e3 * (ta - t0):.2f} ms, expansion {1e3 * (tb - ta):.2f} ms', file=sys.stderr)
        return tb - t0, st, ms
    # first call at this size (pays the allocator's first large blocks), then the reported one
    ewald_first, _, _ = precompute()
    ewald_seconds, est, expand_ms = precompute()
    diff = float((p_chk - p_unit).abs().max() / p_unit.abs().max())
    del p_chk
    ewald_info = {'seconds': round(ewald_seconds, 4), 'seconds_first_call': round(ewald_first, 4), 'k_eff': est['k_eff'],
                  'rows_direct': sc.n_per_cell, 'method': est.get('method'), 'fourier_ms': round(est['fourier_ms'], 3),
                  'finish_ms': round(est['finish_ms'], 3), 'expand_ms': round(expand_ms, 3),
                  'collectives': 'none (every rank evaluates the unit-cell rows and expands the array itself)',
                  'dmma_rows_path': {'seconds': round(rows_seconds, 4), 'fourier_ms': round(est_rows['fourier_ms'], 3),
                                     'allreduce_unit_rows_seconds': round(t_reduce, 4),
                                     'fp64_tflops_fourier': round(4.0 * sc.n_per_cell * N * est_rows['k_eff'] / world
                                                                  / (est_rows['fourier_ms'] * 1e-3) / 1e12, 3),
                                     'fp64_peak_tflops': 37.1,
                                     'max_rel_diff_vs_class_factorised_rows': diff,
                                     'what': 'pycd_ewald_rows(0, n_per_cell) with the k list split over the ranks + '
                                             'one all-reduce: 4 n_per_cell N K_eff flop on the DMMA kernel'}}
    assert diff <= 1e-12, f'class-factorised rows differ from the DMMA rows: {diff:.3e}'

    # ---- KMC system + ensemble -------------------------------------------------------
    if args.dense:
        system = K.KmcSystem(ctx, run, P.data_ptr())
    else:
        system = K.KmcSystem(ctx, run, p_unit.data_ptr(), layout='unit_rows')
    traj_id0 = rank * nt
    occ = K.philox_initial_occupancy(run.tables, nt, C_, seed, traj_id0=traj_id0)
    # time grid: ~1 row per 2 % of the timed KMC steps (k_total ~ C * 3.2e9 /s for Hematite e-)
    total_kmc = (args.steps + args.warmup) * args.kmc_steps
    dt_grid = (total_kmc / (C_ * 3.2e9) * constants.SEC2AUTIME) / max(args.n_path - 1, 1)
    S = args.kmc_steps - (args.kmc_steps % args.refresh if args.refresh > 1 else 0)


    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    kernel_names = {}
    region = {}     # per timed region: when this rank's launches were done, and what the closing barrier added

    def timed_run(refresh, steps, warmup):
        ens = K.KmcEnsemble(system, occ, dt_grid=dt_grid, n_path=args.n_path, step_limit=10 ** 12,
                            stop_at_grid_end=False, rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=traj_id0,
                            refresh_interval=refresh)
        for _ in range(warmup):
            if not args.no_flush:
                ctx.flush_l2()        # also allocates the 512 MB flush buffer outside the timed region
            ens.advance_async(S)
        ens.wait()
        # NVML set-up before the opening barrier: done after it, it delayed rank 0's launches by ~5 ms and every
        # other rank waited that long at the closing barrier
        sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clocks) else None
        barrier()
        ctx.reset_timers()
        l0 = ctx.launch_count()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        # the launches are enqueued back to back (pycd_kmc_advance_async: no host round trip per launch, so
        # host jitter -- other ranks, the NVML sampler -- cannot open gaps between them); wait() + barrier
        # close the timed region
        for _ in range(steps):
            if not args.no_flush:
                ctx.flush_l2()            # 512 MB memset between timed iterations (stream-ordered)
            ens.advance_async(S)
        ens.wait()
        t_done = time.perf_counter() - t0
        barrier()
        wall = time.perf_counter() - t0
        region[refresh] = {'own_launches_done_s': round(t_done, 5), 'closing_barrier_s': round(wall - t_done, 5)}
        clocks = sampler.stop() if sampler else None
        kern_ms = ctx.total_kernel_ms(nat.KC_KMC_STEP)
        launches = ctx.launch_count() - l0
        kernel_names[refresh] = ens.last_kernel()
        return ens, wall, kern_ms, launches, clocks

    ens, wall, kern_ms, launches, clocks = timed_run(args.refresh, args.steps, args.warmup)
    if dist:
        tw = torch.tensor([wall, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall, kern_ms = float(tw[0]), float(tw[1])
    steps_all = world * nt * S * args.steps
    value = steps_all / wall

    # stateless (refresh=1, the B_step formulation) for the roofline comparison
    stateless = None
    if args.refresh != 1:
        ens1, wall1, kern1, _, _ = timed_run(1, max(2, args.steps // 4), 1)
        ens1.close()
        if dist:
            tw = torch.tensor([wall1, kern1], dtype=torch.float64, device=dev)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall1, kern1 = float(tw[0]), float(tw[1])
        n1 = max(2, args.steps // 4)
        stateless = {'value': world * nt * S * n1 / wall1, 'kernel_ms_per_launch': kern1 / n1}

    # ---- e2e: the public API with HOST buffers every step --------------------------------
    e2e_ens = K.KmcEnsemble(system, occ, dt_grid=dt_grid, n_path=args.n_path, step_limit=S,
                            stop_at_grid_end=False, rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=traj_id0,
                            refresh_interval=args.refresh)

    # double-buffered read-back (pycd_kmc_read_begin / _end): the device->host copy of batch k runs on a
    # second stream under the re-arm + launch of batch k+1; every byte still crosses inside the timed region
    e2e_out = [e2e_ens.pinned_buffers(), e2e_ens.pinned_buffers()]
    occ_pinned = nat.pinned_empty(occ.shape, np.int32)
    occ_pinned[...] = occ

    def e2e_run(n):
        # batch i+1 is re-armed and launched BEFORE the host waits for the read-back of batch i: the GPU never
        # idles on the host (reset, advance_async and read_begin only enqueue; read_end is the one wait)
        e2e_ens.reset(occ_pinned, traj_id0)                      # H2D: initial sites (pinned host)
        e2e_ens.advance_async(S)
        for i in range(n):
            e2e_ens.read_begin(e2e_out[i & 1])                   # D2H: displacement grid + state, asynchronous
            if i + 1 < n:
                e2e_ens.reset(occ_pinned, traj_id0)
                e2e_ens.advance_async(S)
            e2e_ens.read_end()                                   # batch i is complete in its host buffers
        e2e_ens.wait()
        return e2e_out[(n - 1) & 1]
    e2e_steps = max(2, min(args.steps, 64))     # the drain of the last read-back is inside the timed region
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    out = e2e_run(e2e_steps)
    barrier()
    e2e_wall = time.perf_counter() - t0
    if dist:
        tw = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e_wall = float(tw[0])
    # the batches are identical (same sites, same Philox keys): the last host buffer must equal the
    # state of a plain read
    e2e_ens.reset(occ_pinned, traj_id0)
    e2e_ens.advance_resident(S)
    chk = e2e_ens.read(unwrapped=True)
    e2e_equal = bool(np.array_equal(chk['unwrapped'], out['unwrapped']) and
                     np.array_equal(chk['occupancy'], out['occupancy']) and
                     np.array_equal(chk['n_steps'], out['n_steps']))
    del chk
    h2d = occ.nbytes
    d2h = out['unwrapped'].nbytes + sum(out[k].nbytes for k in ('n_steps', 'time', 'occupancy', 'drift',
                                                                  'near_tie', 'clamped')) + 4 * nt
    e2e_value = world * nt * S * e2e_steps / e2e_wall

    # ---- MSD on the resident grid + NCCL reduction of the partial sums --------------------
    msd_info = None
    if not args.skip_msd:
        n_msd = args.n_path // 2 + 1
        toff = np.array([0, C_], dtype=np.int32)
        avg = M.species_avg_sd(ctx, ens.unwrapped_device_ptr(), nt, args.n_path, C_, n_msd,
                               1 / constants.ANG2BOHR, toff)
        part = torch.tensor(np.stack([avg[:, :, 0].sum(0), (avg[:, :, 0] ** 2).sum(0)]), device=dev)
        if dist:
            dist.all_reduce(part)
        mean = (part[0] / (world * nt)).cpu().numpy()
        msd_info = {'msd_ms': round(ctx.last_kernel_ms(nat.KC_MSD), 3), 'n_msd': n_msd,
                    'msd_last_A2': float(mean[-1])}
    state = ens.read(unwrapped=True)
    near_tie = int(state['near_tie'].sum())
    ens.close()

    # ---- the other BASELINE configurations as bounded legs (bench_configs.py) ---------------
    configs = None
    if not args.skip_configs and not args.dense:
        import bench_configs as BC
        e2e_ens.close()
        configs = {}
        # cfg 5: 8 conditions x 128 trajectories per GPU (weak: 1024 trajectories per GPU at every N)
        configs['cfg5_sweep'] = BC.cfg5_sweep(ctx, dev, rank, world, dist, system, run,
                                              traj_per_condition=128 * world)
        configs['cfg2_bvo'] = BC.cfg2_bvo(ctx, dev, rank, world, dist)
        p_keep = P if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
        if p_keep is None:
            del P
            torch.cuda.empty_cache()
        configs['cfg4_ewald'] = BC.cfg4_ewald(ctx, dev, rank, world, dist)
        if p_keep is not None:
            P = p_keep

    # ---- CPU baseline on the host cores (rank 0, N=1 only) --------------------------------
    cpu = None
    parity_in_bench = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        P_host = P.cpu().numpy()
        n_t, steps_c, rate1 = size_cpu_sample(run, P_host, occ, args, args.n_path, dt_grid, seed, True)
        rate, dt_c, tot, thr = cpu_reference_rate(run, P_host, occ, args, n_t, steps_c, args.n_path,
                                                  dt_grid, seed, literal=True)
        # the O(C) gather restatement re-runs trajectories 0..n_g-1 of the TIMED ensemble for every KMC step
        # the GPU took (warm-up + timed launches) from the same sites and Philox keys: its final sites,
        # step counts and displacement grids must equal what the GPU holds -- the bench run itself is checked
        total_per_traj = int(state['n_steps'][0])
        n_g = int(max(1, min(host_cores(), len(occ), args.parity_budget // max(total_per_traj, 1))))
        parity = parity_check(run, P_host, occ, state, n_g, total_per_traj, args.n_path, dt_grid, seed, traj_id0)
        rate_g, dt_g, steps_g = parity.pop('rate'), parity.pop('seconds'), total_per_traj
        cpu = {'value': rate, 'unit': 'KMC steps/s', 'cores': thr, 'kind': 'port',
               'sample': f'{n_t} trajectories x {steps_c} steps of the same ensemble, C port of the '
                         f'reference step loop (O(N) dot per process, core.py:2004-2008), '
                         f'{dt_c:.1f} s; 1 core: {rate1:.1f} steps/s',
               'gather_form_value': rate_g,
               'gather_form_sample': f'{n_g} trajectories x {steps_g} steps, O(C) gather restatement, {dt_g:.1f} s'}
        parity_in_bench = parity
        del P_host
        ewald_info['cpu_baseline'] = ewald_cpu_baseline(sc, ep)

    if rank == 0:
        peak, peak_src = measured_peaks()
        bstep = b_step_bytes(n_proc, C_)
        per_launch_bytes = bstep * nt * S
        # bytes the incremental formulation actually gathers per step (8 B elements): untouched
        # carriers 2+2nn each, every carrier nn+1 for the moved carrier's rebuild, + B_step/R
        R_ = max(args.refresh, 1)
        kname = kernel_names.get(args.refresh) or ''
        if 'warp' in kname or 'tpp' in kname:
            # lattice-stencil table: one 8*nn-byte entry per (carrier, other carrier) pair; incremental
            # mode reads 3 entries per carrier per step + the full C x C re-gather every R steps
            touched = 8 * nn * C_ * C_ if R_ == 1 else 8 * nn * (3 * C_ + 2) + 8 * nn * C_ * C_ // R_
        else:
            touched = bstep if R_ == 1 else 8 * ((C_ - 1) * (2 + 2 * nn) + C_ * (nn + 1) + 2 * nn + 2) + bstep // R_
        k_ms = kern_ms / args.steps
        achieved = per_launch_bytes / (k_ms * 1e-3) / 1e9
        floor = step_floor()
        touched_gbs = touched * nt * S / (k_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
        cyc_step = k_ms * 1e-3 * sm_mhz * 1e6 / S
        # The dominant kernel gathers 32-byte table entries that live in L2 (DRAM traffic per launch = one
        # cold fetch of the tables, `traffic`): its bandwidth roofline is the L2 random-sector gather peak
        # measured on this device (tools/step_floor.cu), and what actually bounds it at 3.5 trajectories per
        # SM is the latency of one step's dependency chain (`latency`).
        roofline = {'bound': 'l2', 'achieved': touched_gbs, 'peak': floor['l2_gather_peak_gb_per_s'],
                    'unit': 'GB/s', 'frac': touched_gbs / floor['l2_gather_peak_gb_per_s'],
                    'traffic': profiled_traffic(),
                    'kernel': kernel_names.get(args.refresh), 'kernel_ms_per_launch': k_ms,
                    'touched_bytes_per_kmc_step': touched, 'peak_source': floor['source'],
                    'latency': {'cycles_per_kmc_step': cyc_step, 'floor_cycles': floor['chain_cycles'],
                                'frac': floor['chain_cycles'] / cyc_step, 'floor_terms': floor['chain_terms'],
                                'note': 'one KMC step is ONE dependency chain (rates -> scan -> selection -> '
                                        'gathers -> reduction -> update); floor = sum of the measured '
                                        'dependent-chain latencies of its unavoidable operations'},
                    'algorithmic_ratio': {'bytes_per_kmc_step': bstep, 'algorithmic_bytes_per_launch': per_launch_bytes,
                                          'achieved_gb_per_s': achieved, 'hbm_peak_gb_per_s': peak,
                                          'ratio_to_hbm_peak': achieved / peak, 'hbm_peak_source': peak_src,
                                          'note': 'SURVEY 8(d) algorithmic bytes of the STATELESS gather formulation '
                                                  '(8*n_proc*(2C+6)+4*n_proc per step) x steps/s against the HBM '
                                                  'peak: a ratio above 1 means the table form + incremental sums '
                                                  'read fewer bytes than that formulation counts, not that HBM '
                                                  'is exceeded'}}
        line = {
            'metric': 'KMC steps/s (all trajectories, box-wide)', 'value': value, 'unit': 'KMC steps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': wall / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'Hematite {args.size[0]}x{args.size[1]}x{args.size[2]} supercell '
                                   f'(N={N}), {C_} electrons, {nt} trajectories/GPU '
                                   f'({world * nt} total), Philox draws, fixed-step mode',
                       'kmc_steps_per_step': S, 'n_proc': n_proc, 'refresh_interval': args.refresh, 'p_layout': 'dense N x N' if args.dense else 'unit-cell rows (n_per_cell x N) + lattice translation; step kernel reads the row-difference stencil table built from them',
                       'time_grid_rows': args.n_path,
                       'l2_policy': 'L2 flushed (512 MB write) between timed iterations' + (f'; dense array {N * N * 8 / 1e9:.1f} GB > L2' if args.dense else '; the unit-row table is re-fetched from HBM after each flush'),
                       'parallelism': f'trajectories sharded over {world} GPU(s), no data-path collective'},
            'e2e': {'value': e2e_value, 'unit': 'KMC steps/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps,
                    'read_back_equals_plain_read': e2e_equal,
                    'note': 'per step through the C ABI with pinned host numpy buffers: re-arm the ensemble from '
                            'the host occupancy array, advance, read the displacement grid and state back '
                            '(double-buffered: the copy of batch k runs under batch k+1, pycd_kmc_read_begin/_end); '
                            'the Ewald table stays resident (uploaded once per material, like the reference '
                            'loads precomputed_array.npy once)'},
            'gpu_launches': int(launches), 'timed_region_rank0': region.get(args.refresh),
            'clocks': clocks,
            'roofline': roofline,
            'cpu_baseline': cpu, 'parity_in_bench': parity_in_bench,
            'ewald': ewald_info, 'msd': msd_info, 'near_tie_fallbacks': near_tie,
            'configs': configs,
        }
        if stateless:
            a1 = per_launch_bytes / (stateless['kernel_ms_per_launch'] * 1e-3) / 1e9
            t1 = 8 * nn * C_ * C_ if ('warp' in (kernel_names.get(1) or '')) else bstep
            l2_1 = t1 * nt * S / (stateless['kernel_ms_per_launch'] * 1e-3) / 1e9
            line['stateless'] = {'value': stateless['value'], 'kernel_ms_per_launch': stateless['kernel_ms_per_launch'],
                                 'kernel': kernel_names.get(1), 'touched_bytes_per_kmc_step': t1,
                                 'l2_gather_achieved_gb_per_s': l2_1,
                                 'l2_gather_frac': l2_1 / floor['l2_gather_peak_gb_per_s'],
                                 'algorithmic_gb_per_s': a1, 'algorithmic_ratio_to_hbm_peak': a1 / peak}
        print(json.dumps(line))
        if parity_in_bench and parity_in_bench['equal'] != parity_in_bench['checked']:
            print('bench.py: the timed ensemble differs from the CPU oracle', file=sys.stderr)
            rc = 1
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return rc


# ---------------------------------------------------------------------------------------
def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the step loop (C port of the
    literal O(N)-dot algorithm, oracle/pycd_oracle.c) on all host cores, same config."""
    import torch
    from pycd_b200 import _native as nat
    from pycd_b200 import ewald as EW
    from pycd_b200 import kmc as K
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(0)
    lat, sc, run, ep = build_problem(args)
    N = sc.num_system_elements
    # the array itself is an INPUT of the timed path; the reference cannot build it at this
    # size, so it is produced by the GPU setup code, untimed
    P_host, _ = EW.precomputed_array(ctx, ep)
    return reference_timed(args, run, P_host, N)


def reference_timed(args, run, P_host, N):
    from pycd_b200 import constants
    from pycd_b200 import kmc as K
    C_ = run.n_carriers
    seed = 2
    cores = host_cores()
    occ = K.philox_initial_occupancy(run.tables, max(cores, 1), C_, seed)
    total_kmc = (args.steps + args.warmup) * args.kmc_steps
    dt_grid = (total_kmc / (C_ * 3.2e9) * constants.SEC2AUTIME) / max(args.n_path - 1, 1)
    # bounded sample per step: every core advances one trajectory by `steps_c` KMC steps
    n_t, steps_c, rate1 = size_cpu_sample(run, P_host, occ, args, args.n_path, dt_grid, seed, True)
    budget = max(args.cpu_seconds / max(args.steps + args.warmup, 1), 1.0)
    steps_c = int(max(2, min(steps_c, rate1 * budget)))
    for _ in range(args.warmup):
        cpu_reference_rate(run, P_host, occ, args, n_t, steps_c, args.n_path, dt_grid, seed, literal=True)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        _, _, n, thr = cpu_reference_rate(run, P_host, occ, args, n_t, steps_c, args.n_path, dt_grid, seed,
                                          literal=True)
        tot += n
    wall = time.perf_counter() - t0
    value = tot / wall
    n_proc = C_ * run.tables.nn
    sz = [int(v) for v in run.supercell.system_size]   # the build container (no GPU) times the shipped 2x2x1 cell
    line = {'impl': 'reference', 'metric': 'KMC steps/s (all trajectories, box-wide)', 'value': value,
            'unit': 'KMC steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': wall / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'Hematite {sz[0]}x{sz[1]}x{sz[2]} supercell '
                                   f'(N={N}), {C_} electrons, Philox draws, fixed-step mode',
                       'n_proc': n_proc},
            'cpu_baseline': {'value': value, 'unit': 'KMC steps/s', 'cores': thr, 'kind': 'port',
                             'sample': f'per step: {n_t} trajectories x {steps_c} KMC steps, one trajectory '
                                       'per host thread, C port of the reference step loop (O(N) dot per '
                                       'process); the reference itself is pure Python and cannot travel'},
            'e2e': {'value': value, 'unit': 'KMC steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))
    return 0


def single_config(args):
    """--config 2 / 4 / 5: that BASELINE configuration alone, one JSON line."""
    import torch
    import bench_configs as BC
    from pycd_b200 import _native as nat
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(local_rank)
    sampler = ClockSampler(local_rank).start() if (rank == 0 and not args.no_clocks) else None
    l0 = ctx.launch_count()
    if args.config == 2:
        res = BC.cfg2_bvo(ctx, dev, rank, world, dist, traj_per_gpu=args.traj_per_gpu, launches=max(args.steps, 1))
        metric, hib, scaling = 'KMC steps/s (all trajectories, box-wide)', True, 'weak'
    elif args.config == 4:
        res = BC.cfg4_ewald(ctx, dev, rank, world, dist)
        metric, hib, scaling = 'Ewald precompute s', False, 'strong'
    else:
        from pycd_b200 import ewald as EW, kmc as K
        lat, sc, run, ep = build_problem(args)
        p_unit, _ = EW.unit_cell_rows(ctx, ep, np.ascontiguousarray(sc.coordinates))
        system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
        res = BC.cfg5_sweep(ctx, dev, rank, world, dist, system, run, traj_per_condition=args.sweep_traj,
                            common_grid=args.sweep_common_grid)
        system.close()
        metric, hib, scaling = 'KMC steps/s (all trajectories, box-wide)', True, 'strong'
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line = {'metric': metric, 'value': res['value'], 'unit': res['unit'], 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'higher_is_better': hib, 'scaling': scaling, 'vs_baseline': None,
                'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': res['workload'], 'baseline_config': args.config},
                'gpu_launches': int(ctx.launch_count() - l0), 'clocks': clocks, 'detail': res}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm_small(args):
    """No GPU visible (build container): time the port on the shipped 2x2x1 example."""
    import helpers as H
    ex = H.load_example('hematite', species_count=[min(args.carriers, 8), 0])
    run = H.run_parameters(ex)
    return reference_timed(args, run, ex.P, ex.supercell.num_system_elements)


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        import torch
        if int(os.environ.get('RANK', '0')) != 0:
            sys.exit(0)
        sys.exit(reference_arm(a) if torch.cuda.is_available() else reference_arm_small(a))
    if a.config != 3:
        sys.exit(single_config(a))
    sys.exit(main())
