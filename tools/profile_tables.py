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


def _j(name):
    f = os.path.join(P, name)
    return last_json_line(f) if os.path.exists(f) else None


def _fmt(v):
    return f'{v:.4g}' if v is not None else '–'


def _passed(log):
    f = os.path.join(P, log)
    if not os.path.exists(f):
        return '?'
    m = re.findall(r'(\d+) passed', open(f).read())
    return m[-1] if m else '?'


def readme(tag):
    """The whole of profiles/README.md from the evidence files (python tools/profile_tables.py --readme r02)."""
    head = _j(f'bench_{tag}.json')
    bs64 = _j(f'bench_train_bs64_{tag}.json')
    split = _j(f'bench_infer_split_{tag}.json')
    ref = _j(f'bench_reference_{tag}.json')
    n2, n8 = _j(f'bench_n2_{tag}.json'), _j(f'bench_n8_{tag}.json')
    n4 = _j(f'bench_n4_{tag}.json')
    try:
        pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pk = {}
    out = []
    w = out.append
    w(f'# profiles/ — measured evidence, round {int(tag[1:])}\n')
    w('Produced on B200 boxes through `gpurun` by `tests/run_gpu_round.sh` (one call: the `-m gpu` suite, the kernel-level tests once\n'
      'more under the non-caching allocator, smoke, the bench lines, the per-kernel roofline table, the ncu launch lists and the\n'
      '`ncu --set full` captures) and written up by `tools/profile_tables.py --readme ' + tag + '`.  Round-1 evidence is kept under\n'
      '`profiles/r01/`.  `.ncu-rep` files stay in `gpurun_out/` (scratch); the `ncu_*_' + tag + '.txt` summaries were made from them\n'
      'with `tools/ncu_summary.py`.\n')
    w(f"Peaks (`MEASURED_PEAKS.json`, driver-written): HBM copy {pk.get('hbm_gbs', 0):.0f} GB/s, dense bf16 cuBLAS "
      f"{pk.get('bf16_tflops', 0):.0f} TFLOP/s burst / {pk.get('bf16_tflops_sustained', 0):.1f} TFLOP/s\nsustained.  Tensor-core kernels "
      'are reported against the *sustained* figure (they are timed back to back).  GEMM-shaped kernels are\nreported against the roof '
      'that binds them (the 27 / 70-channel layers have 38-122 FLOP per byte, below the ridge of ~209: HBM).\n')
    if head:
        c = head.get('clocks') or {}
        w(f"The B = 512 step runs into the board power cap (`{', '.join(c.get('reasons', []))}`, SM clock {c.get('sm_mhz', 0):.0f} MHz of "
          f"{c.get('sm_max_mhz', 0):.0f}): run-to-run spread of the headline ±2 %.\n")
    # ---- tests
    w('## Test / parity evidence\n')
    w(f"* `test_gpu_{tag}.log`: {_passed(f'test_gpu_{tag}.log')} GPU tests passed; `test_nocache_{tag}.log`: the "
      f"{_passed(f'test_nocache_{tag}.log')} kernel-level tests again with\n  `PYTORCH_NO_CUDA_MEMORY_CACHING=1`; `smoke_{tag}.log`; "
      f"`sanitizer_{tag}.txt`: compute-sanitizer is closed on this pool, what\n  replaces it.")
    pr = os.path.join(P, f'parity_report_{tag}.jsonl')
    if os.path.exists(pr):
        rows = [json.loads(l) for l in open(pr)]
        w(f'* `parity_report_{tag}.jsonl`: every measured error of the model-level tests.  The discriminative ones (trained-like\n'
          "  full-width reference state, against the REFERENCE's fp32 results):\n")
        w('| fixture | eval max-abs / output range | loss rel. | full-gradient cosine | worst tensor rel. L2 | 20-step Adam trajectory |')
        w('|---|---|---|---|---|---|')
        for v in ('base', 'upr', 'dpp'):
            ev = [r for r in rows if r.get('test') == f'trained_{v}' and r.get('mode') == 'eval' and 'max_abs_of_range' in r]
            tr = [r for r in rows if r.get('test') == f'trained_{v}' and r.get('mode') == 'train']
            tj = [r for r in rows if r.get('test') == f'trajectory_{v}']
            if not (ev and tr and tj):
                continue
            e, t, j = max(ev, key=lambda r: r['max_abs_of_range']), tr[-1], tj[-1]
            w(f"| `net_trained_{v}` | {e['max_abs_of_range']:.2e} | {t['loss_rel']:.1e} | {t['grad_cosine']:.6f} | "
              f"{100 * t['worst_tensor_rel_l2']:.1f} % (`{t['worst_tensor']}`) | worst {100 * j['worst_rel']:.2f} % (loss "
              f"{j['first']:.3f} → {j['last']:.3f}, reference {j['ref_last']:.3f}) |")
        fd = [r for r in rows if r.get('test') == 'finite_difference_trained']
        sc = [r for r in rows if r.get('test') == 'single_activation_copy']
        if fd:
            w(f"\nFinite differences on the trained model: analytic {fd[-1]['analytic']:.3f} vs FD {fd[-1]['fd']:.3f} "
              f"({100 * fd[-1]['rel']:.1f} %; the fp32 reference itself\ngives 0.7-2.9 % at these step sizes)."
              + (f"  Single-activation-copy option vs the default: gradient cosine {sc[-1]['cosine']:.7f}." if sc else ''))
    # ---- bench
    w(f"\n## Bench lines (`bench_{tag}.json` = the driver's default command: headline + `secondary`; N = 1)\n")
    w('| workload | value | e2e (host buffers in, result out, every step) | ms / step | conv roofline in-step (of sustained peak) | step (of sustained peak) |')
    w('|---|---|---|---|---|---|')

    def brow(label, d):
        if not d:
            return
        r = d.get('roofline') or {}
        roof = f"{r['achieved']:.0f} TFLOP/s = {r['frac']:.2f}" if r.get('achieved') else '–'
        if r.get('wgrad'):
            roof += f"; wgrad {r['wgrad']['frac']:.2f}"
            if r['wgrad'].get('wide'):
                roof += f" (280-ch layers {r['wgrad']['wide']['frac']:.2f})"
        bw = r.get('by_layer_width') or {}
        if bw.get('wide') and bw.get('narrow'):
            roof += f"; 280-ch layers {bw['wide']['frac']:.2f}, in-nets {bw['narrow']['frac']:.2f} of the HBM rate"
        stp = f"{r['step']['frac']:.2f}" if r.get('step') else '–'
        w(f"| {label} | {_fmt(d.get('value'))} {d.get('unit', '')} | {_fmt((d.get('e2e') or {}).get('value'))} | "
          f"{d.get('ms_per_step', 0):.2f} | {roof} | {stp} |")
    if head:
        sec = head.get('secondary') or {}
        brow('BASE train bs = 512, 96 px (headline; `TrainStep`, CUDA-graph replay)', head)
        brow('UPR train (`secondary.upr_train`)', sec.get('upr_train'))
        brow('DPP train (`secondary.dpp_train`)', sec.get('dpp_train'))
        brow(f'BASE train, 64 patches = per-GPU share at 8 GPUs (`bench_train_bs64_{tag}.json`)', bs64)
        brow('BASE full-LF inference 512×512 (`secondary.infer`; e2e = 33 uint8 crosshair views in, extraction on the GPU)', sec.get('infer'))
        brow('ESE, 70 members, one LF (`secondary.ese`)', sec.get('ese'))
        brow('row-band sharded inference (`secondary.bands`, N = 1)', sec.get('bands'))
        brow(f'split-precision inference (`bench_infer_split_{tag}.json`)', split)
    if ref:
        cb = ref.get('cpu_baseline') or {}
        w(f"| CPU port of the reference algorithm, {cb.get('cores', '?')} host cores (`bench_reference_{tag}.json`: {cb.get('sample', '')}) | "
          f"{_fmt(ref.get('value'))} {ref.get('unit', '')} | – | {ref.get('ms_per_step', 0):.0f} | – | – |")
    if head:
        km = head.get('kernel_ms_per_step') or {}
        r = head.get('roofline') or {}
        top = ', '.join(f"{k.replace('mmlf_', '')} {v:.1f}" for k, v in list(km.items())[:7])
        w(f"\nHeadline step, per-kernel times of the un-graphed profiling pass (ms; sum {r.get('kernel_sum_ms', 0):.1f}, step "
          f"{head['ms_per_step']:.1f}, idle {100 * (r.get('idle_frac_of_step') or 0):.1f} %, host\nenqueue "
          f"{head.get('host_enqueue_ms_per_step', 0):.3f} ms per step, {head.get('gpu_launches', 0) // max(head.get('steps', 1), 1)} kernel "
          f"launches per step): {top}.\nBatchNorm passes: {100 * r.get('batchnorm_share_of_kernel_time', 0):.1f} % of the kernel time.")
        f32 = ((head.get('secondary') or {}).get('infer') or {}).get('e2e_f32_stacks')
        if f32:
            w(f"Inference e2e with the four float32 stacks as input instead: {f32['value']:.1f} Mpx/s (113 MB of H2D per light field).")
    sa, sa64 = _j(f'bench_train_single_act_{tag}.json'), _j(f'bench_train_bs64_single_act_{tag}.json')
    if sa and sa64:
        w(f"\nSingle-activation-copy option (`MMLF_SINGLE_ACT=1`, `bench_train*_single_act_{tag}.json`, measured earlier in the round): "
          f"bs = 512 {sa['ms_per_step']:.2f} ms, 64 patches {sa64['ms_per_step']:.2f} ms -- see DESIGN.md §4.2.")
    # ---- scaling
    if n2 or n8:
        w(f"\n## Scaling (`bench_n2_{tag}.json`, `bench_n4_{tag}.json`, `bench_n8_{tag}.json`: `torchrun`, one rank per GPU, NCCL all-reduces captured in the step graph)\n")
        w('| workload | N = 1 | N = 2 | N = 4 | N = 8 | 8-GPU speed-up (value / e2e) |')
        w('|---|---|---|---|---|---|')

        def pick(d, key):
            if not d:
                return None
            return d if key is None else (d.get('secondary') or {}).get(key)
        for label, key in (('BASE train, global batch 512', None), ('UPR train', 'upr_train'), ('DPP train', 'dpp_train'),
                           ('full-LF inference, one light field per GPU (weak)', 'infer'), ('ESE, members sharded', 'ese'),
                           ('one light field, row bands', 'bands')):
            a, b, c = pick(head, key), pick(n2, key), pick(n8, key)
            m4 = pick(n4, key)
            sp = '–'
            if a and c:
                sp = f"{c['value'] / a['value']:.2f}× / {c['e2e']['value'] / a['e2e']['value']:.2f}×"
            w(f"| {label} | {_fmt(a and a['value'])} | {_fmt(b and b['value'])} (e2e {_fmt(b and b['e2e']['value'])}) | "
              f"{_fmt(m4 and m4['value'])} (e2e {_fmt(m4 and m4['e2e']['value'])}) | "
              + (f"{_fmt(c['value'])} {c['unit']} ({c['ms_per_step']:.2f} ms)" if c else '–') + f" | {sp} |")
        w(f"\n`multi_check_n2_{tag}.txt`: sharded ESE and row-band inference bit-identical to the single-process result on both ranks; ranks\n"
          'built from different seeds are one replica after the rank-0 broadcast and stay bit-identical over captured training steps.\n'
          f'`multi_check_n4_{tag}.txt`, `multi_check_n8_{tag}.txt`: the same on 4 and 8 ranks.  The row-band line is latency bound (0.8 ms per light field at 8 GPUs).')
    # ---- kernel table
    w(f"\n## Per-kernel roofline (`kernel_roofline_{tag}.jsonl`, `tools/kernel_bench.py`)\n")
    w('Each kernel alone at the BASELINE sizes (64 patches of 96 px = the per-GPU share at 8 GPUs, or one 512×512 light field), 20\n'
      'launches captured in one CUDA graph, the replay timed with CUDA events (round 1 timed eager launches from Python, which for the\n'
      '10-40 µs kernels measured the ctypes call).  `frac` = achieved ÷ measured peak of the roof that binds the kernel (tensor:\n'
      'sustained bf16 rate; HBM: copy rate); the last column gives the other roof.  Rows whose working set fits the 126 MB L2 (`adam`,\n'
      '`lf_extract`, the loss kernels) are L2-assisted in this back-to-back setting, hence fractions above 1; `ese_reduce` is SFU\n'
      'bound (1.28 G Laplace evaluations per launch).\n')
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kernel_table(tag)
    w(buf.getvalue().rstrip())
    w(f"\n## ncu captures (`ncu_*_{tag}.txt`; `--set full --clock-control none`, one launch each, cold cache)\n")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ncu_table(tag)
    w(buf.getvalue().rstrip())
    w(f"\nLaunch lists of the bench commands themselves (`ncu --metrics gpu__time_duration.sum --graph-profiling node`):\n"
      f"`launches_train_bs64_{tag}.csv` (the captured 64-patch training step, node by node), `launches_infer_{tag}.csv`.")
    print('\n'.join(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--bench-tag', default='r01g')
    ap.add_argument('--kernel-tag', default='r01h')
    ap.add_argument('--ncu-tag', default='r01g')
    ap.add_argument('--readme', default=None, metavar='TAG', help='print the whole profiles/README.md for this tag')
    a = ap.parse_args()
    if a.readme:
        readme(a.readme)
        raise SystemExit(0)
    print(f'## bench lines ({a.bench_tag})\n')
    bench_table(a.bench_tag)
    print(f'\n## per-kernel roofline ({a.kernel_tag})\n')
    kernel_table(a.kernel_tag)
    print(f'\n## ncu captures ({a.ncu_tag})\n')
    ncu_table(a.ncu_tag)
