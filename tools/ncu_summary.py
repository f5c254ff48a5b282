#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per captured launch + top stall instructions.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_uniform.sum']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name', '?')}  (ID {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                print(f'   {k:70s} {d[k]:>16s} {u[k]}')
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    for b in blocks[:1]:
        idx = {h: i for i, h in enumerate(b['hdr'])}
        stalls = [h for h in b['hdr'] if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(r[idx['# Samples']]) for r in b['rows'])
        print(f"\n== top {top_n} instructions by warp-stall samples: {b['name']} ({len(b['rows'])} SASS instructions, {tot} samples)")
        agg = {}
        for r in b['rows']:
            for h in stalls:
                agg[h] = agg.get(h, 0) + int(r[idx[h]])
        print('   stall totals: ' + ', '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
        for r in sorted(b['rows'], key=lambda r: -int(r[idx['# Samples']]))[:top_n]:
            s = int(r[idx['# Samples']])
            st = sorted(((h[6:], int(r[idx[h]])) for h in stalls if int(r[idx[h]])), key=lambda kv: -kv[1])[:3]
            print(f"   {s:6d} {100.0 * s / max(tot, 1):5.1f}%  exec={r[idx['Instructions Executed']]:>9s}  {r[1].strip()[:72]:72s} {st}")


if __name__ == '__main__':
    main()
