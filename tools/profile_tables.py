#!/usr/bin/env python
"""Print the markdown tables of profiles/README.md from the files next to it (bench lines, per-kernel roofline, ncu
summaries), so the numbers in the text can be regenerated instead of typed.
usage: tools/profile_tables.py [--bench-tag r01g] [--kernel-tag r01h] [--ncu-tag r01g]"""
import argparse
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')


def last_json_line(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith('{'):
            return json.loads(line)
    return None


def bench_table(tag):
    print('| file | value | e2e (host buffers) | ms / step | conv roofline in-step (of sustained peak) | SM MHz |')
    print('|---|---|---|---|---|---|')
    for f in sorted(glob.glob(os.path.join(P, f'bench_*_{tag}.json'))):
        d = last_json_line(f)
        if not d:
            continue
        r = d.get('roofline') or {}
        clk = (d.get('clocks') or {}).get('sm_mhz')
        roof = f"{r['achieved']:.0f} {r['unit']} = {r['frac']:.2f}" if r.get('achieved') else '–'
        print(f"| `{os.path.basename(f)}` | {d['value']:.4g} {d['unit']} | {d['e2e']['value']:.4g} | {d['ms_per_step']:.2f} | "
              f"{roof} | {clk if clk else '–'} |")


def kernel_table(tag):
    print('| kernel | case | time | bound | achieved | frac of that roof | other roof |')
    print('|---|---|---|---|---|---|---|')
    for line in open(os.path.join(P, f'kernel_roofline_{tag}.jsonl')):
        d = json.loads(line)
        other = ''
        if d.get('frac_of_tensor_peak') is not None:
            other = f"{d['tflops']:.0f} TFLOP/s = {d['frac_of_tensor_peak']:.2f} of tensor"
        elif d.get('frac_of_hbm_peak'):
            other = f"{d['frac_of_hbm_peak']:.2f} of HBM"
        print(f"| {d['kernel'].replace('_kernel', '')} | {d['case']} | {d['ms'] * 1e3:.1f} µs | {d['bound']} | "
              f"{d['achieved']:.0f} {d['unit']} | {d['frac']:.2f} | {other} |")


def ncu_table(tag):
    def metric(path, key):
        for line in open(path):
            parts = line.split()
            if len(parts) >= 2 and parts[0] == key:
                return float(parts[1]), (parts[2] if len(parts) > 2 else '')
        return None, ''
    print('| file | kernel | time | tensor pipe (elapsed) | DRAM read + written |')
    print('|---|---|---|---|---|')
    for f in sorted(glob.glob(os.path.join(P, f'ncu_*_{tag}.txt'))):
        name = re.sub(r'\(.*', '', open(f).readline().replace('== ', '').replace('void ', '')).strip()
        t, tu = metric(f, 'gpu__time_duration.sum')
        tp, _ = metric(f, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')
        r, ru = metric(f, 'dram__bytes_read.sum')
        w, wu = metric(f, 'dram__bytes_write.sum')
        print(f"| `{os.path.basename(f)}` | {name} | {t:.1f} {tu} | {tp:.1f} % | {r:.1f} {ru} + {w:.1f} {wu} |")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--bench-tag', default='r01g')
    ap.add_argument('--kernel-tag', default='r01h')
    ap.add_argument('--ncu-tag', default='r01g')
    a = ap.parse_args()
    print(f'## bench lines ({a.bench_tag})\n')
    bench_table(a.bench_tag)
    print(f'\n## per-kernel roofline ({a.kernel_tag})\n')
    kernel_table(a.kernel_tag)
    print(f'\n## ncu captures ({a.ncu_tag})\n')
    ncu_table(a.ncu_tag)
