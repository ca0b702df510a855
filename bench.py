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
    ap.add_argument('--skip-msd', action='store_true')
    ap.add_argument('--no-flush', action='store_true', help='diagnostic: skip the L2 flush between iterations')
    ap.add_argument('--no-clocks', action='store_true', help='diagnostic: do not sample nvidia-smi clocks')
    ap.add_argument('--dense', action='store_true', help='step kernel gathers from the dense N x N array')
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
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

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
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    p_unit = torch.empty((sc.n_per_cell, N), dtype=torch.float64, device=dev)
    _, est = EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=p_unit.data_ptr())
    EW.ewald_expand(ctx, sc, p_unit.data_ptr(), r0, r1, out=P[r0:r1].data_ptr())
    expand_ms = ctx.last_kernel_ms(nat.KC_EWALD_EXPAND)
    torch.cuda.synchronize()
    t_gather = 0.0
    if dist:
        tg = time.perf_counter()
        if N % world == 0:
            dist.all_gather_into_tensor(P.view(-1), P[r0:r1].reshape(-1).clone())
        else:
            for g in range(world):
                a, b = (g * N) // world, ((g + 1) * N) // world
                dist.broadcast(P[a:b], src=g)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - tg
    ewald_seconds = time.perf_counter() - t0
    ewald_info = {'seconds': round(ewald_seconds, 4), 'k_eff': est['k_eff'], 'rows_direct': sc.n_per_cell,
                  'fourier_ms': round(est['fourier_ms'], 3), 'finish_ms': round(est['finish_ms'], 3),
                  'expand_ms': round(expand_ms, 3), 'allgather_seconds': round(t_gather, 4),
                  'fp64_tflops_fourier': round(4.0 * sc.n_per_cell * N * est['k_eff'] / (est['fourier_ms'] * 1e-3) / 1e12, 3)
                  if est['fourier_ms'] > 0 else None,
                  'method': 'rows of unit cell 0 on the GPU + translation expansion of the rank\'s row block'}

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

    def timed_run(refresh, steps, warmup):
        ens = K.KmcEnsemble(system, occ, dt_grid=dt_grid, n_path=args.n_path, step_limit=10 ** 12,
                            stop_at_grid_end=False, rng_mode=nat.RNG_PHILOX, seed=seed, traj_id0=traj_id0,
                            refresh_interval=refresh)
        for _ in range(warmup):
            ens.advance_resident(S)
        barrier()
        ctx.reset_timers()
        l0 = ctx.launch_count()
        sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clocks) else None
        t0 = time.perf_counter()
        for _ in range(steps):
            if not args.no_flush:
                ctx.flush_l2()            # 512 MB memset between timed iterations
            ens.advance_resident(S)
        barrier()
        wall = time.perf_counter() - t0
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
        for i in range(n):
            e2e_ens.reset(occ_pinned, traj_id0)                  # H2D: initial sites (pinned host)
            e2e_ens.advance_resident(S)
            if i > 0:
                e2e_ens.read_end()                               # batch i-1 is complete in its host buffers
            e2e_ens.read_begin(e2e_out[i & 1])                   # D2H: displacement grid + state, asynchronous
        e2e_ens.read_end()
        return e2e_out[(n - 1) & 1]
    e2e_steps = max(2, min(args.steps, 8))
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
    state = ens.read(unwrapped=False)
    near_tie = int(state['near_tie'].sum())
    ens.close()

    # ---- CPU baseline on the host cores (rank 0, N=1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        P_host = P.cpu().numpy()
        n_t, steps_c, rate1 = size_cpu_sample(run, P_host, occ, args, args.n_path, dt_grid, seed, True)
        rate, dt_c, tot, thr = cpu_reference_rate(run, P_host, occ, args, n_t, steps_c, args.n_path,
                                                  dt_grid, seed, literal=True)
        n_g, steps_g, rate1g = size_cpu_sample(run, P_host, occ, args, args.n_path, dt_grid, seed, False)
        rate_g, dt_g, tot_g, _ = cpu_reference_rate(run, P_host, occ, args, n_g, steps_g, args.n_path,
                                                    dt_grid, seed, literal=False)
        cpu = {'value': rate, 'unit': 'KMC steps/s', 'cores': thr, 'kind': 'port',
               'sample': f'{n_t} trajectories x {steps_c} steps of the same ensemble, C port of the '
                         f'reference step loop (O(N) dot per process, core.py:2004-2008), '
                         f'{dt_c:.1f} s; 1 core: {rate1:.1f} steps/s',
               'gather_form_value': rate_g,
               'gather_form_sample': f'{n_g} trajectories x {steps_g} steps, O(C) gather restatement, {dt_g:.1f} s'}
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
                    'd2h_bytes_per_step': int(d2h),
                    'read_back_equals_plain_read': e2e_equal,
                    'note': 'per step through the C ABI with pinned host numpy buffers: re-arm the ensemble from '
                            'the host occupancy array, advance, read the displacement grid and state back '
                            '(double-buffered: the copy of batch k runs under batch k+1, pycd_kmc_read_begin/_end); '
                            'the Ewald table stays resident (uploaded once per material, like the reference '
                            'loads precomputed_array.npy once)'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': profiled_traffic(),
                         'kernel': kernel_names.get(args.refresh),
                         'kernel_ms_per_launch': k_ms,
                         'algorithmic_bytes_per_launch': per_launch_bytes,
                         'bytes_per_kmc_step': bstep, 'peak_source': peak_src,
                         'touched_bytes_per_kmc_step': touched,
                         'touched_achieved': touched * nt * S / (k_ms * 1e-3) / 1e9,
                         'note': 'algorithmic bytes are those of the STATELESS gather formulation (SURVEY 8d: '
                                 '8*n_proc*(2C+6)+4*n_proc per KMC step); the kernel reads a precomputed '
                                 'row-difference table (one 32-byte entry per carrier pair) and, with '
                                 'refresh_interval>1, patches cached sums, so it touches touched_bytes_per_kmc_step '
                                 'and frac exceeds 1; the step is bound by the latency of its dependency chain '
                                 '(DESIGN.md 4.3), see stateless for the like-for-like figure'},
            'cpu_baseline': cpu,
            'ewald': ewald_info, 'msd': msd_info, 'near_tie_fallbacks': near_tie,
        }
        if stateless:
            a1 = per_launch_bytes / (stateless['kernel_ms_per_launch'] * 1e-3) / 1e9
            line['stateless'] = {'value': stateless['value'], 'kernel_ms_per_launch': stateless['kernel_ms_per_launch'],
                                 'roofline_achieved': a1, 'roofline_frac': a1 / peak, 'kernel': kernel_names.get(1)}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    line = {'impl': 'reference', 'metric': 'KMC steps/s (all trajectories, box-wide)', 'value': value,
            'unit': 'KMC steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': wall / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'Hematite {args.size[0]}x{args.size[1]}x{args.size[2]} supercell '
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
    sys.exit(main())
