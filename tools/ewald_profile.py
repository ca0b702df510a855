#!/usr/bin/env python
"""One Ewald precompute of the benchmark system for profiling / A-B timing of the fp64
reciprocal-space kernel: the rows of unit cell 0 (skinny 32x256 tiles) and, with --dense ROWS,
a dense row block (64x128 tiles).  Prints one JSON line with the kernel times.

    python tools/ewald_profile.py [--size 10 10 10] [--dense 256] [--repeat 3]
    ncu --set full --import-source on --clock-control none -k regex:ewald_fourier -c 1 \\
        -o gpurun_out/ewald python tools/ewald_profile.py --repeat 1
"""
import argparse
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, nargs=3, default=[10, 10, 10])
    ap.add_argument('--example', default='hematite')
    ap.add_argument('--dense', type=int, default=0, help='also time a dense block of this many rows')
    ap.add_argument('--repeat', type=int, default=3)
    args = ap.parse_args()
    import yaml
    from pycd_b200 import _native as nat
    from pycd_b200 import ewald as EW
    from pycd_b200.lattice import Lattice, Supercell
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(0)
    d = ROOT / 'tests' / 'golden' / args.example
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    lat = Lattice(SimpleNamespace(**cfg))
    sc = Supercell(lat, args.size, [1, 1, 1])
    ep = EW.EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    coords = np.ascontiguousarray(sc.coordinates)
    N = sc.num_system_elements
    out = {'N': int(N), 'size': args.size}
    unit, dense = [], []
    for _ in range(args.repeat):
        p_unit, st = EW.ewald_rows(ctx, ep, coords, 0, sc.n_per_cell)
        unit.append(st['fourier_ms'])
    out.update(k_eff=int(st['k_eff']), unit_rows=int(sc.n_per_cell), unit_ms=unit,
               unit_tflops=4.0 * sc.n_per_cell * N * st['k_eff'] / (min(unit) * 1e-3) / 1e12)
    if args.dense:
        r0 = N // 2
        for _ in range(args.repeat):
            blk, st = EW.ewald_rows(ctx, ep, coords, r0, r0 + args.dense)
            dense.append(st['fourier_ms'])
        ref = EW.ewald_expand(ctx, sc, p_unit, r0, r0 + args.dense)
        out.update(dense_rows=args.dense, dense_ms=dense,
                   dense_tflops=4.0 * args.dense * N * st['k_eff'] / (min(dense) * 1e-3) / 1e12,
                   dense_vs_expanded=float(np.abs(blk - ref).max() / np.abs(ref).max()))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
