#!/usr/bin/env python3
"""SASS of the production step kernel's hot loop, with the mnemonics DESIGN.md 4.3 talks about counted.

usage: sass_excerpt.py <libpycd_b200.so> <out.txt>
Takes kmc_step_warp_kernel<2,1,4,INCR=true,MODE=0 (plain)> (the benchmark shape) out of `cuobjdump -sass`, finds
the burst loop (the last backward branch of the function that spans the DMMAs) and writes that range."""
import re
import subprocess
import sys

FUN = '_ZN4pycd20kmc_step_warp_kernelILi2ELi1ELi4ELb1ELi0EEEvNS_6SysDevENS_10StencilDevENS_6EnsDevENS_11AdvanceArgsE'


def main():
    lib, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', FUN, lib], capture_output=True, text=True).stdout
    ins = []
    for ln in txt.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    dmma = [i for i, (_, t) in enumerate(ins) if t.startswith('DMMA')]
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA(?:\.U)?\s+.*?(0x[0-9a-f]+)', t)
        if m and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            if j < i and dmma and j <= dmma[0] and i >= dmma[-1]:
                if best is None or (i - j) < (best[1] - best[0]):
                    best = (j, i)
    lo, hi = best
    body = ins[lo:hi + 1]
    count = lambda pat: sum(1 for _, t in body if re.search(pat, t))
    with open(out, 'w') as f:
        f.write(f'# cuobjdump -sass -fun kmc_step_warp_kernel<2,1,4,INCR,PLAIN> {lib}\n')
        f.write(f'# burst loop: {len(body)} instructions, 0x{body[0][0]:04x} .. 0x{body[-1][0]:04x} (cold blocks included)\n')
        for name, pat in (('DMMA.8x8x4 (warp scan: 4, direction sums: 5)', r'^DMMA'), ('LDG.E.NA.ENL2.256.CONSTANT (table entries, no L1 allocation)', r'LDG\.E\.(NA\.)?ENL2\.256'),
                          ('REDUX.SUM (selection count)', r'^REDUX'), ('SHFL', r'SHFL'), ('BAR.SYNC', r'BAR\.SYNC'),
                          ('BSSY', r'^BSSY'), ('VOTE', r'^VOTE'), ('DFMA/DMUL/DADD', r'^D(FMA|MUL|ADD)'), ('MUFU.RCP64H (time advance division)', r'MUFU\.RCP64H'),
                          ('LDS', r'^(@!?U?P\d\s+)?LDS'), ('STS', r'^(@!?U?P\d\s+)?STS')):
            f.write(f'#   {count(pat):4d}  {name}\n')
        for a, t in body:
            f.write(f'/*{a:04x}*/  {t} ;\n')
    print(out, len(body), 'instructions')


if __name__ == '__main__':
    main()
