#!/usr/bin/env python3
"""Extracts the judged metrics of an .ncu-rep (ncu --set full) into a small JSON + stall table.

usage: summarize_ncu.py <report.ncu-rep> <out.json> [kernel-substring-for-line-profile cubin]
"""
import csv
import io
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {'kernel': r[hdr.index('Kernel Name')]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    k[w] = float(r[i].replace(',', ''))
                except ValueError:
                    k[w] = r[i]
                k[w + ' unit'] = units[i]
        kernels.append(k)
    src = ncu_csv(rep, 'source')
    sh = src[1]
    cols = [(i, h) for i, h in enumerate(sh) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {}
    for r in src[2:]:
        for i, h in cols:
            try:
                tot[h] = tot.get(h, 0) + (int(r[i] or 0) if i < len(r) else 0)
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    stalls = {h: round(100 * v / s, 2) for h, v in sorted(tot.items(), key=lambda x: -x[1]) if v}
    json.dump({'report': rep.split('/')[-1], 'kernels': kernels, 'warp_stall_samples_pct': stalls},
              open(dst, 'w'), indent=1)
    print(json.dumps(kernels[0], indent=1)[:1500])
    print(stalls)


if __name__ == '__main__':
    main()
