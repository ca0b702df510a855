#!/usr/bin/env python
"""BASELINE config 4: "BVO large supercell Ewald precompute only (N^2 site-pair array,
row-sharded + NVLink all-gather)".

Each rank evaluates its row block of the dense N x N array DIRECTLY (no translation
symmetry: 4*N^2*K_eff/G flop per GPU), the blocks are all-gathered over NCCL, and the result
is checked at full size against the translation-expanded unit-cell rows (a size-independent
property: max|P_dense - P_expanded| <= 1e-12 * max|P|), plus symmetry P = P^T.

    python tools/ewald_cfg4.py [--size 12 12 6]
    torchrun --nproc-per-node N tools/ewald_cfg4.py --size 12 12 6
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, nargs=3, default=[12, 12, 6])
    ap.add_argument('--example', default='bvo')
    ap.add_argument('--max-rows', type=int, default=0, help='only the first MAX_ROWS rows of each block (sampling)')
    args = ap.parse_args()
    import torch
    import yaml
    from pycd_b200 import _native as nat
    from pycd_b200 import dist as D
    from pycd_b200 import ewald as EW
    from pycd_b200.lattice import Lattice, Supercell
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
        os.dup2(saved, 1)
    if nat.needs_build():
        nat.build()
    ctx = nat.default_context(local)
    d = ROOT / 'tests' / 'golden' / args.example
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    lat = Lattice(SimpleNamespace(**cfg))
    sc = Supercell(lat, args.size, [1, 1, 1])
    ep = EW.EwaldParameters(sc, cfg['alpha'], cfg['r_cut'], cfg['k_cut'])
    n = sc.num_system_elements
    lo, hi = D.block(rank, world, n)
    if args.max_rows:
        hi = min(hi, lo + args.max_rows)
    P = torch.zeros((n, n), dtype=torch.float64, device=dev)
    coords = torch.from_numpy(np.ascontiguousarray(sc.coordinates)).to(dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    _, st = EW.ewald_rows(ctx, ep, coords.data_ptr(), lo, hi, out=P[lo:hi].data_ptr())
    t_kernel = time.perf_counter() - t0
    t_gather = 0.0
    if world > 1 and not args.max_rows:
        tg = time.perf_counter()
        D.allgather_rows(P)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - tg
    total = time.perf_counter() - t0
    # full-size check against the translation-expanded unit-cell rows
    pu = torch.empty((sc.n_per_cell, n), dtype=torch.float64, device=dev)
    EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=pu.data_ptr())
    chk_lo, chk_hi = (0, n) if (world > 1 and not args.max_rows) or world == 1 and not args.max_rows else (lo, hi)
    err, scale, asym = 0.0, float(pu.abs().max()), 0.0
    step = 2048
    for r0 in range(chk_lo, chk_hi, step):
        r1 = min(chk_hi, r0 + step)
        ref = torch.empty((r1 - r0, n), dtype=torch.float64, device=dev)
        EW.ewald_expand(ctx, sc, pu.data_ptr(), r0, r1, out=ref.data_ptr())
        torch.cuda.synchronize()
        err = max(err, float((P[r0:r1] - ref).abs().max()))
        if chk_lo == 0 and chk_hi == n:
            asym = max(asym, float((P[r0:r1] - P[:, r0:r1].T).abs().max()))
    if world > 1:
        tt = torch.tensor([t_kernel, total, err, asym], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_kernel, total, err, asym = (float(v) for v in tt)
    if rank == 0:
        flops = 4.0 * (hi - lo) * n * st['k_eff']
        print(json.dumps({
            'config': f'{args.example} {args.size} dense Ewald, rows sharded over {world} GPU(s)',
            'n_sites': n, 'k_eff': st['k_eff'], 'rows_per_gpu': hi - lo, 'k_split': st['k_split'],
            'seconds_total': round(total, 4), 'seconds_fourier_kernel': round(st['fourier_ms'] / 1e3, 4),
            'seconds_finish_kernel': round(st['finish_ms'] / 1e3, 4), 'seconds_allgather': round(t_gather, 4),
            'fp64_tflops_per_gpu': round(flops / (st['fourier_ms'] * 1e-3) / 1e12, 2),
            'max_abs_diff_vs_translation_expanded': err, 'max_abs_P': scale, 'rel': err / scale,
            'max_asymmetry': asym}))
        assert err <= 1e-12 * scale, 'dense rows disagree with the translation-expanded rows'
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
