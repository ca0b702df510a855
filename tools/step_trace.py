#!/usr/bin/env python
"""Per-phase cycle counts of one KMC step of the stencil kernel (diagnostic).

Needs a library built with PYCD_NVCC_EXTRA=-DPYCD_TRACE: the kernel then stamps clock64() at
the phase boundaries of trajectory 0 for the first 256 steps of a launch.  Usage (GPU box):
    PYCD_NVCC_EXTRA=-DPYCD_TRACE python -c "from pycd_b200 import _native as n; n.build(force=True)"
    python tools/step_trace.py [--traj 512] [--refresh 256]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

NAMES_TPP = ['top', 'rates', 'k stored', 'after (1)', 'selected', 'loads issued', 'reduced', 'before C', 'after C', 'end']
NAMES_WARP = ['top', 'rates', 'args ready', 'exp done', 'k stored', 'prefix done', 'scan done', 'selected',
              'owner loads issued', 'time/rows done', 'loads landed', 'sums updated', 'end', 'lookups done',
              'H loads issued', 'time advanced']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--traj', type=int, default=512)
    ap.add_argument('--refresh', type=int, default=256)
    ap.add_argument('--steps', type=int, default=1024)
    a = ap.parse_args()
    import bench
    from pycd_b200 import _native as nat, ewald as EW, kmc as K
    args = bench.parse_args.__wrapped__() if hasattr(bench.parse_args, '__wrapped__') else None
    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    ctx = nat.default_context(0)
    lat, sc, run, ep = bench.build_problem(args)
    p_unit, _ = EW.ewald_rows(ctx, ep, np.ascontiguousarray(sc.coordinates), 0, sc.n_per_cell)
    system = K.KmcSystem(ctx, run, p_unit, layout='unit_rows')
    occ = K.philox_initial_occupancy(run.tables, a.traj, run.n_carriers, 2)
    ens = K.KmcEnsemble(system, occ, dt_grid=1e9, n_path=11, step_limit=10 ** 12, stop_at_grid_end=False,
                        rng_mode=nat.RNG_PHILOX, seed=2, refresh_interval=a.refresh)
    ens.advance_resident(a.steps)
    ens.advance_resident(a.steps)
    print(ens.last_kernel())
    warp_kernel = 'warp' in ens.last_kernel()
    NAMES = NAMES_WARP if warp_kernel else NAMES_TPP
    out = np.zeros(256 * 16 * 16, dtype=np.int64)
    fn = nat.lib().pycd_debug_trace
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_int
    nat.check(fn(out.ctypes.data))
    tr = out.reshape(256, 16, 16)
    n_warps = 1 if warp_kernel else 9
    steps = slice(8, 250)
    for w in ([0, 1] if warp_kernel else [0, 3, 7, 8]):
        t = tr[steps, w, :16].astype(np.float64)
        t0 = tr[steps, 0, 0].astype(np.float64)[:, None]
        print(f'warp {w}: mean cycles since the step top of warp 0')
        rel = t - t0
        order = np.argsort(np.where((tr[steps, w, :len(NAMES)] > 0).all(axis=0), rel[:, :len(NAMES)].mean(axis=0), 1e18))
        for i in order:
            nme = NAMES[i]
            col = rel[:, i]
            ok = tr[steps, w, i] > 0
            if ok.any():
                print(f'   {i:2d} {nme:14s} {col[ok].mean():9.1f}  (min {col[ok].min():8.0f}, max {col[ok].max():8.0f})')
    per_step = np.diff(tr[8:250, 0, 0]).astype(np.float64)
    print('cycles per step (warp 0 top to top): mean %.1f  median %.1f' % (per_step.mean(), np.median(per_step)))


if __name__ == '__main__':
    main()
