#!/usr/bin/env python
"""Per-kernel roofline table of the hot path (GPU box): every kernel of include/mmlf_b200.h at the BASELINE.json sizes,
timed with CUDA events over back-to-back launches (the stream stays full, so host launch cost is hidden), against the
measured peaks of MEASURED_PEAKS.json.  One JSON line per kernel on stdout; `--only NAME[|NAME...]` restricts to kernels whose name
contains one of the NAMEs (used as the ncu target: `ncu --set full -k regex:... python tools/kernel_bench.py --only lf_shift --reps 1`).

Algorithmic bytes / flops per unit follow SURVEY.md section 8(d) and DESIGN.md section 4.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mmlf_b200 import _lib, ops  # noqa: E402
from mmlf_b200._lib import BF16, FP16, ConvArgs, call  # noqa: E402

DEV = 'cuda'
TD = {BF16: torch.bfloat16, FP16: torch.float16}


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ST():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad16(x):
    return (x + 15) // 16 * 16


def timeit(fn, reps, warm=3, allow_graph=True):
    """Seconds per launch.  The `reps` launches are captured into ONE CUDA graph and the replay is timed: the C-ABI is
    driven from Python through ctypes (15-40 us of host time per call), so back-to-back eager launches of a 20-40 us kernel
    measure the host, not the kernel (round 1 reported `adam` at 35 us and `lf_extract` at 44 us; ncu and the graph replay
    agree on 19 and 24 us).  Falls back to eager launches when a call cannot be captured."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if allow_graph and os.environ.get('MMLF_BENCH_EAGER', '0') != '1' and reps > 1:
        try:
            graph = torch.cuda.CUDAGraph()
            with _lib.no_gc_during_capture(), torch.cuda.graph(graph):
                for _ in range(reps):
                    fn()
            graph.replay()                      # warm replay
            torch.cuda.synchronize()
        except Exception:
            graph = None
            torch.cuda.synchronize()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for _ in range(reps):
            fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


class Bench:
    def __init__(self, args):
        self.args = args
        try:
            pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pk = {}
        self.hbm = pk.get('hbm_gbs', 6500.0)
        self.tf_burst = pk.get('bf16_tflops', 1600.0)
        self.tf_sus = pk.get('bf16_tflops_sustained', 1400.0)
        self.rows = []

    def want(self, name):
        return self.args.only is None or any(o in name for o in self.args.only.split('|'))

    def hbm_row(self, name, what, nbytes, fn, note='', allow_graph=True):
        if not self.want(name + ' ' + what):
            return
        t = timeit(fn, self.args.reps, allow_graph=allow_graph)
        gbs = nbytes / t / 1e9
        row = {'kernel': name, 'case': what, 'bound': 'hbm', 'ms': t * 1e3, 'algorithmic_MB': nbytes / 1e6,
               'achieved': gbs, 'peak': self.hbm, 'unit': 'GB/s', 'frac': gbs / self.hbm, 'note': note}
        self.rows.append(row)
        print(json.dumps(row), flush=True)

    def tc_row(self, name, what, flops, fn, note='', nbytes=0):
        """A GEMM-shaped kernel: its roof is the larger of flops / sustained tensor peak and algorithmic bytes / HBM rate
        (the 27- and 70-channel layers have 67-122 FLOP per byte, below the ridge of ~209: they are HBM-bound)."""
        if not self.want(name + ' ' + what):
            return
        t = timeit(fn, self.args.reps)
        tf = flops / t / 1e12
        gbs = nbytes / t / 1e9
        f_tc, f_hbm = tf / self.tf_sus, gbs / self.hbm
        row = {'kernel': name, 'case': what, 'ms': t * 1e3, 'algorithmic_GFLOP': flops / 1e9,
               'algorithmic_MB': nbytes / 1e6, 'flop_per_byte': flops / nbytes if nbytes else None}
        if f_hbm > f_tc:
            row.update(bound='hbm', achieved=gbs, peak=self.hbm, unit='GB/s', frac=f_hbm, frac_of_tensor_peak=f_tc,
                       tflops=tf)
        else:
            row.update(bound='tensor', achieved=tf, peak=self.tf_sus, unit='TFLOP/s', frac=f_tc,
                       frac_of_burst_peak=tf / self.tf_burst, frac_of_hbm_peak=f_hbm)
        row['note'] = note
        self.rows.append(row)
        print(json.dumps(row), flush=True)


def conv_case(bn, B, H, W, cin, cout, ctype, dt=FP16, relu=True, dual=False, bits=False, stats=False, bnbwd=False):
    n_slots = B * (H + 1) * (W + 1)
    cin_pad, n_pad = pad16(cin), pad16(cout)
    x = (torch.randn((n_slots, cin_pad), device=DEV) * 0.5).to(TD[dt])
    kc = (cin_pad + 63) // 64
    w = torch.from_numpy(np.random.RandomState(0).normal(0, 0.03, (cout, cin, 2, 2)).astype(np.float32)).to(DEV)
    wp = torch.empty((n_pad, 4 * kc * 64), dtype=TD[dt], device=DEV)
    call('mmlf_pack_conv_weight', P(w), cout, cin, 0, 0, 1, cin, cin_pad, P(wp), n_pad, cin_pad, dt, ST())
    out = torch.empty((n_slots, n_pad), dtype=TD[dt], device=DEV)
    out2 = torch.empty((n_slots, n_pad), dtype=torch.bfloat16, device=DEV) if dual else None
    rb = torch.empty((n_slots, (n_pad + 31) // 32), dtype=torch.int32, device=DEV) if bits else None
    bias = torch.zeros(n_pad, device=DEV)
    sums = torch.zeros(2 * n_pad, dtype=torch.float64, device=DEV) if (stats or bnbwd) else None
    zbn = torch.randn((n_slots, n_pad), device=DEV).to(torch.float16) if bnbwd else None
    bnc = torch.rand((3, n_pad), device=DEV) if bnbwd else None
    a = ConvArgs()
    a.in_, a.ld_in, a.cin_pad, a.wpack, a.n_pad = x.data_ptr(), cin_pad, cin_pad, wp.data_ptr(), n_pad
    a.B, a.H, a.W, a.type = B, H, W, ctype
    a.bias, a.relu = bias.data_ptr(), int(relu)
    a.relu_bits = rb.data_ptr() if bits else None
    a.ld_bits = (n_pad + 31) // 32
    a.out, a.ld_out, a.out_mode = out.data_ptr(), n_pad, 0
    a.out2 = out2.data_ptr() if dual else None
    a.ld_out2 = n_pad
    a.col_sums = sums.data_ptr() if sums is not None else None
    if bnbwd:
        a.bn_z, a.ld_z, a.bn_z_dtype = zbn.data_ptr(), n_pad, FP16
        a.bn_scale, a.bn_shift, a.bn_mean = bnc[0].data_ptr(), bnc[1].data_ptr(), bnc[2].data_ptr()
    a.ab_dtype, a.out_dtype, a.out2_dtype = dt, dt, BF16
    keep = (x, wp, out, out2, rb, bias, sums, w, zbn, bnc)
    m = n_slots if ctype == 0 else B * H * W
    flops = 2.0 * m * cout * 4 * cin
    # algorithmic HBM bytes: the input slots once, every output array once (weights: < 1 MB, L2 resident)
    nbytes = n_slots * 2 * (cin_pad + n_pad * (2 if dual else 1) + (n_pad if bnbwd else 0)) + \
        (n_slots * 4 * ((n_pad + 31) // 32) if bits else 0)

    def fn():
        call('mmlf_conv2x2', C.byref(a), ST())
    fn.keep = keep
    return fn, flops, nbytes


def wgrad_case(B, H, W, cin, cout, ctype, act_dt=BF16):
    n_slots = B * (H + 1) * (W + 1)
    cin_pad, n_pad = pad16(cin), pad16(cout)
    act = (torch.randn((n_slots, cin_pad), device=DEV) * 0.5).to(TD[act_dt])
    dout = (torch.randn((n_slots, n_pad), device=DEV) * 0.5).to(torch.bfloat16)
    ws = torch.empty(_lib.lib().mmlf_conv2x2_wgrad_workspace(n_pad, cin_pad) // 4, dtype=torch.float32, device=DEV)
    dw = torch.empty(n_pad * 4 * cin_pad, dtype=torch.float32, device=DEV)
    m = n_slots if ctype == 0 else B * H * W

    def fn():
        call('mmlf_conv2x2_wgrad', P(dout), n_pad, n_pad, P(act), cin_pad, cin_pad, B, H, W, ctype, act_dt, BF16, P(ws),
             P(dw), ST())
    fn.keep = (act, dout, ws, dw)
    return fn, 2.0 * m * cout * 4 * cin, n_slots * 2 * (cin_pad + n_pad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--only', default=None)
    ap.add_argument('--train-batch', type=int, default=64, help='patches per GPU (512 / 8 GPUs)')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    _lib.require_device()
    bn = Bench(args)
    Bt, ps = args.train_batch, 96
    g = torch.Generator(device=DEV).manual_seed(0)

    # ------------------------------------------------------------------ tensor-core kernels
    for name, (B, H, W, cin, cout, ct, kw) in {
        'conv2x2 280->280 pad1 train': (Bt, ps, ps, 280, 280, 0, dict(dual=True, bits=True)),
        'conv2x2 280->280 pad0 train': (Bt, ps, ps, 280, 280, 1, dict(relu=False, stats=True)),
        'conv2x2 280->280 pad0 dgrad': (Bt, ps, ps, 280, 280, 1, dict(dt=BF16, relu=False)),
        'conv2x2 280->280 pad0 dgrad + BN-bwd statistics': (Bt, ps, ps, 280, 280, 1, dict(dt=BF16, relu=False, bnbwd=True)),
        'conv2x2 70->70 pad0 dgrad': (Bt, ps, ps, 70, 70, 1, dict(dt=BF16, relu=False)),
        'conv2x2 70->70 pad0 dgrad + BN-bwd statistics': (Bt, ps, ps, 70, 70, 1, dict(dt=BF16, relu=False, bnbwd=True)),
        'conv2x2 70->70 pad1 train': (Bt, ps, ps, 70, 70, 0, dict(dual=True, bits=True)),
        'conv2x2 70->70 pad0 train': (Bt, ps, ps, 70, 70, 1, dict(relu=False, stats=True)),
        'conv2x2 27->70 pad1 train': (Bt, ps, ps, 27, 70, 0, dict(dual=True, bits=True)),
        'conv2x2 280->280 pad1 infer': (1, 512, 512, 280, 280, 0, {}),
        'conv2x2 280->280 pad0 infer': (1, 512, 512, 280, 280, 1, {}),
        'conv2x2 70->70 pad0 infer': (1, 512, 512, 70, 70, 1, {}),
        'conv2x2 280->108 pad1 infer': (1, 512, 512, 280, 108, 0, {}),
    }.items():
        if bn.want('conv2x2_tc2_kernel ' + name):
            fn, flops, nbytes = conv_case(bn, B, H, W, cin, cout, ct, **kw)
            bn.tc_row('conv2x2_tc2_kernel', name, flops, fn, nbytes=nbytes)
            del fn
            torch.cuda.empty_cache()
    for name, (B, H, W, cin, cout, ct) in {
        'wgrad 280->280 pad1 train': (Bt, ps, ps, 280, 280, 0),
        'wgrad 280->280 pad0 train': (Bt, ps, ps, 280, 280, 1),
        'wgrad 70->70 pad0 train': (Bt, ps, ps, 70, 70, 1),
        'wgrad 27->70 pad1 train': (Bt, ps, ps, 27, 70, 0),
    }.items():
        if bn.want('conv2x2_wgrad2_kernel+wgrad_reduce_kernel ' + name):
            fn, flops, nbytes = wgrad_case(B, H, W, cin, cout, ct)
            bn.tc_row('conv2x2_wgrad2_kernel+wgrad_reduce_kernel', name, flops, fn, nbytes=nbytes)
        mixed = name + ', fp16 act x bf16 grad'
        if bn.want('conv2x2_wgrad2_kernel+wgrad_reduce_kernel ' + mixed):
            fn, flops, nbytes = wgrad_case(B, H, W, cin, cout, ct, FP16)
            bn.tc_row('conv2x2_wgrad2_kernel+wgrad_reduce_kernel', mixed, flops, fn,
                      'the MMLF_SINGLE_ACT=1 option: activation boxes converted to bf16 in shared memory', nbytes=nbytes)
            del fn
            torch.cuda.empty_cache()

    # ------------------------------------------------------------------ BatchNorm passes (slot arrays, 16-bit)
    for C_real in (280, 70):
        Cp = pad16(C_real)
        geo_slots = Bt * (ps + 1) * (ps + 1)
        px = Bt * ps * ps
        z = torch.randn((geo_slots, Cp), device=DEV).to(torch.float16)
        dy = torch.randn((geo_slots, Cp), device=DEV).to(torch.bfloat16)
        y = torch.empty_like(z)
        y2 = torch.empty((geo_slots, Cp), dtype=torch.bfloat16, device=DEV)
        dz = torch.empty_like(y2)
        scale = torch.rand(Cp, device=DEV) + 0.5
        shift = torch.randn(Cp, device=DEV) * 0.1
        mean = torch.zeros(Cp, device=DEV)
        invstd = torch.ones(Cp, device=DEV)
        gamma = torch.ones(Cp, device=DEV)
        sums = torch.zeros(2 * Cp, dtype=torch.float64, device=DEV)
        fsums = torch.empty(3 * Cp, device=DEV)
        dgam, dbet, db2 = torch.empty(C_real, device=DEV), torch.empty(C_real, device=DEV), torch.zeros(Cp, device=DEV)
        bn.hbm_row('slot_map_kernel<bn_apply_relu>', f'C={C_real} train, fp16 z -> fp16 y + bf16 y', px * C_real * 6.0,
                   lambda: call('mmlf_bn_apply_relu', P(z), Cp, P(scale), P(shift), Cp, Bt, ps, ps, FP16, P(y), Cp, P(y2), Cp,
                                BF16, ST()), 'what the training step launches: read z, write y twice (conv operand fp16 + wgrad operand bf16)')
        bn.hbm_row('slot_map_kernel<bn_apply_relu>', f'C={C_real} train, fp16 z -> fp16 y', px * C_real * 4.0,
                   lambda: call('mmlf_bn_apply_relu', P(z), Cp, P(scale), P(shift), Cp, Bt, ps, ps, FP16, P(y), Cp, P(None), Cp,
                                BF16, ST()), 'the MMLF_SINGLE_ACT=1 option: read z, write y once')
        bn.hbm_row('col_reduce_kernel<bn_bwd_reduce>', f'C={C_real} train', px * C_real * 4.0,
                   lambda: call('mmlf_bn_bwd_reduce', P(dy), Cp, P(z), Cp, P(scale), P(shift), P(mean), P(invstd), Cp, Bt, ps,
                                ps, BF16, FP16, P(sums), ST()), 'read dy, z')
        bn.hbm_row('slot_map_kernel<bn_bwd_apply>', f'C={C_real} train', px * C_real * 6.0,
                   lambda: call('mmlf_bn_bwd_apply', P(dy), Cp, P(z), Cp, P(scale), P(shift), P(gamma), P(mean), P(invstd),
                                P(sums), px, 1, C_real, Cp, Bt, ps, ps, BF16, FP16, P(dz), Cp, P(dgam), P(dbet), 0, P(fsums),
                                P(db2), ST()), 'read dy, z, write dz (+ column sums of dz)')
        del z, dy, y, y2, dz
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ light-field resampling / packing
    for tag, (B, H, W) in {'full LF 512x512': (1, 512, 512), f'train batch {Bt}x96x96': (Bt, ps, ps)}.items():
        views = [torch.rand((B, 9, 3, H, W), device=DEV, generator=g) for _ in range(4)]
        outs = [torch.empty_like(v) for v in views]
        nel = 4 * B * 27 * H * W
        bn.hbm_row('lf_shift_kernel', f'{tag}, disp 2.5, 4 stacks', nel * 8.0,
                   lambda: call('mmlf_lf_shift', *[P(v) for v in views], *[P(o) for o in outs], B, 9, H, W, 2.5, ST()),
                   'read once + write once, f32')
        slots = torch.empty((B * (H + 1) * (W + 1), 32), dtype=torch.float16, device=DEV)
        bn.hbm_row('pack_views_kernel', f'{tag}, one stack', B * 27 * H * W * 4.0 + B * H * W * 64.0,
                   lambda: call('mmlf_pack_views', P(views[0]), B, 27, H, W, P(slots), 32, FP16, ST()),
                   'read f32 planes, write 32-channel fp16 slots')
        bn.hbm_row('shift_pack_kernel', f'{tag}, stack i (two lerps), disp 2.5', B * 27 * H * W * 4.0 + B * H * W * 64.0,
                   lambda: call('mmlf_shift_pack', P(views[2]), 2, B, 9, H, W, 2.5, P(slots), 32, FP16, ST()),
                   'Shift fused into the packing')
        s4 = [torch.empty((B * (H + 1) * (W + 1), 32), dtype=torch.float16, device=DEV) for _ in range(4)]
        s4g = [torch.empty((B * (H + 1) * (W + 1), 32), dtype=torch.bfloat16, device=DEV) for _ in range(4)]
        PA, IA = C.c_void_p * 4, C.c_int * 4
        vp, sp, gp, ia = PA(*[v.data_ptr() for v in views]), PA(*[t.data_ptr() for t in s4]), \
            PA(*[t.data_ptr() for t in s4g]), IA(0, 1, 2, 3)
        bn.hbm_row('pack_views_kernel (pack_stacks)', f'{tag}, 4 stacks in one launch', 4 * (B * 27 * H * W * 4.0 + B * H * W * 64.0),
                   lambda: call('mmlf_pack_stacks', vp, ia, 4, B, 9, H, W, sp, None, 32, FP16, BF16, 0, 0.0, ST()),
                   'what an inference forward launches: read f32 planes, write 32-channel fp16 slots')
        bn.hbm_row('pack_views_kernel (pack_stacks)', f'{tag}, 4 stacks, fp16 + bf16 copies',
                   4 * (B * 27 * H * W * 4.0 + 2 * B * H * W * 64.0),
                   lambda: call('mmlf_pack_stacks', vp, ia, 4, B, 9, H, W, sp, gp, 32, FP16, BF16, 0, 0.0, ST()),
                   'what a training forward launches: one read, both formats written')
        bn.hbm_row('shift_pack_kernel (pack_stacks)', f'{tag}, 4 stacks in one launch, disp 2.5',
                   4 * (B * 27 * H * W * 4.0 + B * H * W * 64.0),
                   lambda: call('mmlf_pack_stacks', vp, ia, 4, B, 9, H, W, sp, None, 32, FP16, BF16, 1, 2.5, ST()),
                   'what an ESE member launches: Shift fused into the packing')
        del views, outs, slots, s4, s4g
        torch.cuda.empty_cache()
    u8 = torch.randint(0, 256, (81, 512, 512, 3), dtype=torch.uint8, device=DEV, generator=g)
    bn.hbm_row('lf_extract_kernel', '81 u8 views 512x512 -> 4 stacks + centre f32',
               33 * 512 * 512 * 3 * 1.0 + 37 * 3 * 512 * 512 * 4.0, lambda: ops.lf_extract_u8(u8, 9),
               '33 distinct views read, 36 + 1 f32 planes written')
    del u8

    # ------------------------------------------------------------------ augmentation chain
    if bn.want('augment_views_kernel+augment_contrast_kernel'):
        import random
        from mmlf_b200.data.augment import GpuAugmenter
        rs = np.random.RandomState(0)
        scenes = []
        for _ in range(4):
            st = [rs.uniform(0, 1, (9, 3, 512, 512)).astype(np.float32) for _ in range(4)]
            gt_ = rs.uniform(-2, 2, (512, 512)).astype(np.float32)
            mp = np.zeros((1, 5, 512, 512), np.float32)
            scenes.append((*st, st[1][4], gt_, mp, np.ones((512, 512), np.int32), np.atleast_1d(0)))
        aug = GpuAugmenter(scenes)
        ids, params = aug.draw(Bt, ps, random.Random(0), 4)
        out_bytes = Bt * (4 * 27 + 3) * ps * ps * 4.0
        # algorithmic: every output element written once (+ read and re-written by Contrast) and ~1 source element read
        packed = aug.pack(ids, params)
        bn.hbm_row('augment_views_kernel+augment_contrast_kernel', f'{Bt} patches of 96 px from 512x512 scenes, full chain',
                   4.0 * out_bytes, lambda: aug.run(packed, Bt, ps),
                   'gather with stride f (down-sampling) + 2-4 lerp taps; the two kernels over packed sample records')
        bn.hbm_row('GpuAugmenter.__call__', f'{Bt} patches of 96 px: host packing of the records + upload + the two kernels',
                   4.0 * out_bytes, lambda: aug(ids, params),
                   'host bound (ctypes record packing); the training loop prepares batch i + 1 while step i runs',
                   allow_graph=False)

    # ------------------------------------------------------------------ heads, targets, losses, ESE reduce, Adam
    B, H, W, S = Bt, ps, ps, 108
    px = B * H * W
    mean = torch.randn((B, H, W), device=DEV, generator=g)
    logvar = torch.randn((B, H, W), device=DEV, generator=g) * 0.3
    gt = torch.randn((B, H, W), device=DEV, generator=g)
    mask = (torch.rand((B, H, W), device=DEV, generator=g) > 0.2).to(torch.int32)
    scores = torch.randn((B, S, H, W), device=DEV, generator=g)
    bins_t, bins_n = ops.torch_bins(-3.5, 3.5, S, DEV), ops.numpy_bins(-3.5, 3.5, S, DEV)
    sums = ops.loss_prepass(mask)
    bn.hbm_row('loss_prepass_kernel', f'mask {B}x96x96', px * 4.0, lambda: ops.loss_prepass(mask))
    bn.hbm_row('loss_regression_kernel<UPR>', f'{B}x96x96 value + gradients', px * 24.0,
               lambda: ops.loss_regression(2, mean, logvar, gt, mask, None, sums), 'small (14 MB): launch / latency bound')
    bn.hbm_row('loss_regression_kernel<L1>', f'{B}x96x96 value + gradient', px * 16.0,
               lambda: ops.loss_regression(0, mean, None, gt, mask, None, sums))
    bn.hbm_row('loss_ce_kernel', f'{B}x108x96x96 on-the-fly target, value + gradient', px * (8.0 * S + 12),
               lambda: ops.loss_cross_entropy(scores, None, gt, bins_t, 3.5 / S, mask, sums))
    tgt = ops.reg_to_class_op(gt, bins_t, 3.5 / S)
    bn.hbm_row('loss_ce_kernel', f'{B}x108x96x96 dense target, value + gradient', px * (12.0 * S + 8),
               lambda: ops.loss_cross_entropy(scores, tgt, None, None, 0.0, mask, sums))
    bn.hbm_row('reg_to_class_kernel', f'{B}x96x96 -> 108 bins', px * (4.0 + 4 * S), lambda: ops.reg_to_class_op(gt, bins_t, 3.5 / S))
    bn.hbm_row('upr_posterior_kernel', f'{B}x96x96 -> 108 bins', px * (8.0 + 4 * S), lambda: ops.upr_posterior(mean, logvar, bins_n))
    bn.hbm_row('dpp_head_kernel', f'{B}x108x96x96', px * (4.0 * S + 8.0 * S + 8), lambda: ops.dpp_head(scores, bins_t, bins_n),
               'scores read once (shared-memory tile), one_hot + posterior + mean + logvar written')
    del scores, tgt
    K = 70
    means = torch.randn((K, 1, 512, 512), device=DEV, generator=g)
    logvars = torch.randn((K, 1, 512, 512), device=DEV, generator=g) * 0.3
    disp = ops.numpy_bins(-3.5, 3.5, K, DEV)
    if bn.want('ese_reduce'):
        t = timeit(lambda: ops.ese_reduce(means, logvars, disp), args.reps)
        nb = 512 * 512 * (8.0 * K + 4 * K + 8)
        evals = 512 * 512 * K * K
        row = {'kernel': 'ese_reduce_kernel', 'case': '70 members 512x512', 'bound': 'sfu', 'ms': t * 1e3,
               'algorithmic_MB': nb / 1e6, 'achieved': nb / t / 1e9, 'peak': bn.hbm, 'unit': 'GB/s', 'frac': nb / t / 1e9 / bn.hbm,
               'note': f'{evals / 1e9:.2f} G Laplace evaluations (exp) per launch = {evals / t / 1e12:.2f} Texp/s: SFU bound, '
                       f'not HBM bound'}
        print(json.dumps(row), flush=True)
        bn.rows.append(row)
    n = 4612166
    p_, g_, m_, v_ = (torch.randn(n, device=DEV) for _ in range(4))
    v_.abs_()
    bn.hbm_row('adam_kernel', '4,612,166 parameters', n * 28.0, lambda: ops.adam_step(p_, g_, m_, v_, 1e-3, 0.9, 0.999, 1e-8, 3),
               '18 MB working set: L2 resident, launch / latency bound')
    if args.out:
        with open(args.out, 'w') as f:
            for r in bn.rows:
                f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
