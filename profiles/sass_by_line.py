#!/usr/bin/env python3
"""Aggregates an `ncu --page source --csv` SASS listing by CUDA source line.

usage: sass_by_line.py <source_page.csv> <kernel-substring> <cubin> [top]
Line info comes from `nvdisasm -g` of the cubin (compile with -lineinfo); instructions are
matched by position inside the kernel.
"""
import collections
import csv
import re
import subprocess
import sys


def line_map(cubin, kernel_sub):
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    lines = out.splitlines()
    res, cur, active = [], None, False
    for ln in lines:
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel_sub in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+', ln):
            res.append(cur)
    return res


def main():
    page, ksub, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(page)))
    hdr = rows[1]
    i_s, i_e = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    lm = line_map(cubin, ksub)
    data = rows[2:]
    if len(lm) != len(data):
        print(f'warning: {len(lm)} instructions in the cubin vs {len(data)} in the profile', file=sys.stderr)
    samp, exe = collections.Counter(), collections.Counter()
    for k, r in enumerate(data):
        key = lm[k] if k < len(lm) and lm[k] else ('?', 0)
        samp[key] += int(r[i_s] or 0)
        exe[key] += int(r[i_e] or 0)
    ts, te = sum(samp.values()), sum(exe.values())
    print(f'total samples {ts}, warp instructions {te}')
    print('  samples%   instr%   file:line')
    for key, v in samp.most_common(top):
        print(f'  {100 * v / ts:7.2f}  {100 * exe[key] / te:7.2f}   {key[0]}:{key[1]}')


if __name__ == '__main__':
    main()
