#!/usr/bin/env python
"""A/B builds of the lattice-stencil step kernel (diagnostic).

    python tools/step_ab.py build            # here: cross-compiles the variant libraries (scratch/ab/*.so)
    python tools/step_ab.py run [--traj 512 148]   # on the GPU box: bench.py kernel time per variant

A variant = kmc_stencil_nn4.cu recompiled with -D toggles of kmc_stencil.cuh, linked with the objects of
the default build (pycd_b200/csrc/build/libpycd_b200/)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / 'scratch' / 'ab'

VARIANTS = {
    'default': [],
    'shuffle_scan_and_sums': ['-DPYCD_SCAN_DMMA=0', '-DPYCD_SUM_DMMA=0'],
    'helper_warp_off': ['-DPYCD_HELPER_WARP=0'],
    'flat_tail_off': ['-DPYCD_FLAT_TAIL=0'],
    'owner_late': ['-DPYCD_OWNER_EARLY=0'],
    'exp_table': ['-DPYCD_EXP_TABLE=1'],
    'perm_global_load': ['-DPYCD_PERM_PREFETCH=0'],
    'ld_l1_allocate': ['-DPYCD_LD_MODE=0'],
    'ld_cg': ['-DPYCD_LD_MODE=5'],
}
# a header from the history compiled against today's kmc_types.cuh: HEADER@<git rev>
HISTORY = {}


def build():
    from pycd_b200 import _native as nat
    nat.build(force=False)
    objs = [nat.OBJ_DIR / 'libpycd_b200' / (Path(s).stem + '.o') for s in nat.SOURCES]
    OUT.mkdir(parents=True, exist_ok=True)
    env = dict(os.environ)
    env.pop('CC', None)
    env.pop('CXX', None)
    nvcc = nat._nvcc()
    procs = []
    for name, flags in VARIANTS.items():
        obj = OUT / f'nn4_{name}.o'
        cmd = [nvcc] + nat.NVCC_FLAGS + flags + ['-c', '-o', str(obj), str(nat.CSRC_DIR / 'kmc_stencil_nn4.cu')]
        procs.append((name, obj, subprocess.Popen(cmd, env=env)))
    for name, rev in HISTORY.items():
        src_dir = OUT / f'src_{name}'
        src_dir.mkdir(exist_ok=True)
        inc = OUT.parent / 'include'   # common.cuh includes ../../include/pycd_b200.h
        inc.mkdir(exist_ok=True)
        (inc / 'pycd_b200.h').write_bytes((ROOT / 'include' / 'pycd_b200.h').read_bytes())
        for f in nat.CSRC_DIR.iterdir():
            if f.suffix in ('.cu', '.cuh', '.h'):
                (src_dir / f.name).write_bytes(f.read_bytes())
        old = subprocess.run(['git', 'show', f'{rev}:pycd_b200/csrc/kmc_stencil.cuh'], capture_output=True, cwd=ROOT, check=True)
        (src_dir / 'kmc_stencil.cuh').write_bytes(old.stdout)
        obj = OUT / f'nn4_{name}.o'
        cmd = [nvcc] + nat.NVCC_FLAGS + ['-I', str(ROOT / 'include'), '-c', '-o', str(obj), str(src_dir / 'kmc_stencil_nn4.cu')]
        procs.append((name, obj, subprocess.Popen(cmd, env=env)))
    for name, obj, p in procs:
        assert p.wait() == 0, name
        link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(OUT / f'lib_{name}.so')] + \
               [str(obj if o.name == 'kmc_stencil_nn4.o' else o) for o in objs]
        subprocess.run(link, check=True, env=env)
        print('built', name)


def run(trajs, repeats=3):
    rows = []
    names = [n for n in list(VARIANTS) + list(HISTORY) if (OUT / f'lib_{n}.so').exists()]
    for rep in range(repeats):
      for name in names:
        lib = OUT / f'lib_{name}.so'
        for nt in trajs:
            env = dict(os.environ, PYCD_B200_LIB=str(lib))
            res = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--steps', '5', '--warmup', '3',
                                  '--no-cpu-baseline', '--skip-msd', '--skip-configs', '--traj-per-gpu', str(nt)],
                                 env=env, capture_output=True, text=True)
            if res.returncode != 0:
                print(name, nt, 'FAILED', res.stderr[-500:])
                continue
            d = json.loads(res.stdout.strip().splitlines()[-1])
            r = d['roofline']
            row = {'variant': name, 'rep': rep, 'traj': nt, 'kernel_ms': r['kernel_ms_per_launch'],
                   'stateless_ms': (d.get('stateless') or {}).get('kernel_ms_per_launch'),
                   'cycles_per_step': r['latency']['cycles_per_kmc_step'], 'msteps_per_s': d['value'] / 1e6}
            rows.append(row)
            print(json.dumps(row), flush=True)
    (ROOT / 'gpurun_out').mkdir(exist_ok=True)
    json.dump(rows, open(ROOT / 'gpurun_out' / 'step_ab.json', 'w'), indent=1)
    print('---- median kernel ms per launch')
    import statistics
    for name in names:
        for nt in trajs:
            v = [r['kernel_ms'] for r in rows if r['variant'] == name and r['traj'] == nt]
            if v:
                sl = [r['stateless_ms'] for r in rows if r['variant'] == name and r['traj'] == nt and r['stateless_ms']]
                print(f'{name:44s} traj {nt:4d}  median {statistics.median(v):.3f}  min {min(v):.3f}  ({len(v)} runs)'
                      + (f'  stateless {statistics.median(sl):.2f}' if sl else ''))


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build()
    else:
        trajs = [int(v) for v in sys.argv[3:]] if len(sys.argv) > 3 else [512, 148]
        run(trajs)
