"""Out-of-bounds WRITE detection without compute-sanitizer (-m gpu).  compute-sanitizer is closed on this GPU pool
("runs under it have left GPUs needing a reset", profiles/sanitizer_r02.txt), so every output buffer of the kernels below is
carved out of a larger allocation whose surrounding guard bands hold a canary pattern; after the launch the guards must be
untouched and the payload fully written.  Shapes are ragged on purpose (sizes that are not multiples of the tile / vector
widths).  Out-of-bounds READS cannot be seen this way; they are covered by the parity tests on the same ragged shapes
(a stray read changes a result) and by running the suite with PYTORCH_NO_CUDA_MEMORY_CACHING=1 once per round
(tests/run_gpu_round.sh), where a read past a cudaMalloc'd buffer faults."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096            # elements on each side


class Guarded:
    def __init__(self, shape, dtype, fill=None):
        n = int(np.prod(shape))
        self.raw = torch.empty(n + 2 * GUARD, dtype=dtype, device='cuda')
        self.canary = 77 if not dtype.is_floating_point else -1234.5
        self.raw.fill_(self.canary)
        self.t = self.raw[GUARD:GUARD + n].view(shape)
        if fill is not None:
            self.t.copy_(fill)
        self.n = n

    def check(self, what):
        torch.cuda.synchronize()
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.n:]
        assert bool((lo == self.canary).all()), f'{what}: wrote BEFORE the buffer'
        assert bool((hi == self.canary).all()), f'{what}: wrote PAST the buffer'


def _u():
    import _gpu_util as u
    return u


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.mark.parametrize('case', [(1, 5, 7, 27, 70, 0), (2, 9, 11, 70, 70, 1), (1, 13, 6, 280, 280, 0), (1, 7, 9, 280, 280, 1),
                                  (1, 6, 6, 280, 2, 0), (1, 5, 8, 280, 108, 0), (3, 3, 3, 108, 108, 1)])
def test_conv_and_wgrad_stay_inside_their_buffers(case):
    u = _u()
    B, H, W, cin, cout, ctype = case
    rng = np.random.RandomState(1)
    Hp, Wp = H + 1, W + 1
    n_slots = B * Hp * Wp
    cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
    x = rng.normal(0, 1, (B, H, W, cin) if ctype == 0 else (B, Hp, Wp, cin)).astype(np.float32)
    xs = u.to_slots(x, cin_pad, ctype == 1, Hp, Wp, u.FP16)
    w = rng.normal(0, 0.2, (cout, cin, 2, 2)).astype(np.float32)
    wp = u.pack_weight(w, dt=u.FP16)
    out = Guarded((n_slots, n_pad), torch.float16)
    out2 = Guarded((n_slots, n_pad), torch.bfloat16)
    bits = Guarded((n_slots, (n_pad + 31) // 32), torch.int32)
    sums = Guarded((2 * n_pad,), torch.float64, fill=torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda'))
    a = u.ConvArgs()
    a.in_, a.ld_in, a.cin_pad, a.wpack, a.n_pad = xs.data_ptr(), cin_pad, cin_pad, wp.data_ptr(), n_pad
    a.B, a.H, a.W, a.type, a.relu = B, H, W, ctype, 1
    a.out, a.ld_out, a.out_mode = out.t.data_ptr(), n_pad, 0
    a.out2, a.ld_out2 = out2.t.data_ptr(), n_pad
    a.relu_bits, a.ld_bits = bits.t.data_ptr(), (n_pad + 31) // 32
    a.col_sums = sums.t.data_ptr()
    a.ab_dtype, a.out_dtype, a.out2_dtype = u.FP16, u.FP16, u.BF16
    u.call('mmlf_conv2x2', C.byref(a), u.stream())
    for g, nm in ((out, 'conv out'), (out2, 'conv out2'), (bits, 'conv relu bits'), (sums, 'conv col sums')):
        g.check(nm)
    assert bool(torch.isfinite(out.t.float()).all())
    # weight gradient: workspace and canonical gradient
    ws_bytes = u._lib.lib().mmlf_conv2x2_wgrad_workspace(n_pad, cin_pad)
    ws = Guarded((ws_bytes // 4,), torch.float32)
    dw = Guarded((cout, cin, 2, 2), torch.float32)
    gout = u.to_slots(rng.normal(0, 1, (B, Hp, Wp, cout) if ctype == 0 else (B, H, W, cout)).astype(np.float32), n_pad,
                      ctype == 0, Hp, Wp, u.BF16)
    xb = xs.to(torch.bfloat16)
    u.call('mmlf_conv2x2_wgrad_canonical', u.ptr(gout), n_pad, n_pad, u.ptr(xb), cin_pad, cin_pad, B, H, W, ctype, u.BF16,
           u.BF16, _p(ws.t), cout, cin, 0, 1, cin, cin_pad, _p(dw.t), 0, u.stream())
    ws.check('wgrad workspace')
    dw.check('wgrad canonical gradient')
    assert bool(torch.isfinite(dw.t).all()) and bool((dw.t != dw.canary).all())


def test_slot_passes_and_light_field_kernels_stay_inside_their_buffers():
    u = _u()
    from mmlf_b200 import ops
    B, H, W, Cr = 2, 7, 9, 70
    Cp = u.pad16(Cr)
    n_slots = B * (H + 1) * (W + 1)
    z = (torch.randn((n_slots, Cp), device='cuda')).to(torch.float16)
    scale, shift = torch.rand(Cp, device='cuda') + 0.5, torch.randn(Cp, device='cuda')
    y, y2 = Guarded((n_slots, Cp), torch.float16), Guarded((n_slots, Cp), torch.bfloat16)
    u.call('mmlf_bn_apply_relu', u.ptr(z), Cp, u.ptr(scale), u.ptr(shift), Cp, B, H, W, u.FP16, _p(y.t), Cp, _p(y2.t), Cp,
           u.BF16, u.stream())
    y.check('bn_apply_relu y'), y2.check('bn_apply_relu y2')
    gy = torch.randn((n_slots, Cp), device='cuda').to(torch.bfloat16)
    mean, invstd = torch.randn(Cp, device='cuda') * 0.1, torch.rand(Cp, device='cuda') + 0.5
    sums = Guarded((2 * Cp,), torch.float64, fill=torch.zeros(2 * Cp, dtype=torch.float64, device='cuda'))
    u.call('mmlf_bn_bwd_reduce', u.ptr(gy), Cp, u.ptr(z), Cp, u.ptr(scale), u.ptr(shift), u.ptr(mean), u.ptr(invstd), Cp, B, H,
           W, u.BF16, u.FP16, _p(sums.t), u.stream())
    sums.check('bn_bwd_reduce sums')
    dz, fs = Guarded((n_slots, Cp), torch.bfloat16), Guarded((3 * Cp,), torch.float32)
    dgam, dbet = Guarded((Cr,), torch.float32), Guarded((Cr,), torch.float32)
    dzs = Guarded((Cp,), torch.float32, fill=torch.zeros(Cp, device='cuda'))
    gamma = torch.rand(Cr, device='cuda') + 0.5
    u.call('mmlf_bn_bwd_apply', u.ptr(gy), Cp, u.ptr(z), Cp, u.ptr(scale), u.ptr(shift), u.ptr(gamma), u.ptr(mean),
           u.ptr(invstd), _p(sums.t), B * H * W, 1, Cr, Cp, B, H, W, u.BF16, u.FP16, _p(dz.t), Cp, _p(dgam.t), _p(dbet.t), 0,
           _p(fs.t), _p(dzs.t), u.stream())
    for g, nm in ((dz, 'dz'), (fs, 'fsums'), (dgam, 'dgamma'), (dbet, 'dbeta'), (dzs, 'dz colsum')):
        g.check('bn_bwd_apply ' + nm)
    # light-field kernels on a width that is not a multiple of 4 and one that is
    for (Hh, Ww) in ((10, 13), (12, 16)):
        views = [torch.rand((B, 9, 3, Hh, Ww), device='cuda') for _ in range(4)]
        outs = [Guarded((B, 9, 3, Hh, Ww), torch.float32) for _ in range(4)]
        u.call('mmlf_lf_shift', *[u.ptr(v) for v in views], *[_p(o.t) for o in outs], B, 9, Hh, Ww, 2.5, u.stream())
        [o.check('lf_shift') for o in outs]
        ns = B * (Hh + 1) * (Ww + 1)
        s16 = [Guarded((ns, 32), torch.float16) for _ in range(4)]
        s16b = [Guarded((ns, 32), torch.bfloat16) for _ in range(4)]
        PA, IA = C.c_void_p * 4, C.c_int * 4
        for do_shift in (0, 1):
            u.call('mmlf_pack_stacks', PA(*[v.data_ptr() for v in views]), IA(0, 1, 2, 3), 4, B, 9, Hh, Ww,
                   PA(*[g.t.data_ptr() for g in s16]), PA(*[g.t.data_ptr() for g in s16b]), 32, u.FP16, u.BF16, do_shift, -1.3,
                   u.stream())
            [g.check('pack_stacks') for g in s16 + s16b]
    u8 = torch.randint(0, 256, (81, 8, 12, 3), dtype=torch.uint8, device='cuda')
    st = [Guarded((9, 3, 8, 12), torch.float32) for _ in range(4)]
    cen = Guarded((3, 8, 12), torch.float32)
    u.call('mmlf_lf_extract_u8', u.ptr(u8), 9, 8, 12, *[_p(g.t) for g in st], _p(cen.t), u.stream())
    [g.check('lf_extract') for g in st + [cen]]
    # heads / losses / ESE reduce / Adam
    Bq, S, HW = 2, 108, 7 * 9
    scores = torch.randn((Bq, S, 7, 9), device='cuda')
    bt, bnn = ops.torch_bins(-3.5, 3.5, S, 'cuda'), ops.numpy_bins(-3.5, 3.5, S, 'cuda')
    oh, po = Guarded((Bq, S, 7, 9), torch.float32), Guarded((Bq, S, 7, 9), torch.float32)
    mm, lv = Guarded((Bq, 7, 9), torch.float32), Guarded((Bq, 7, 9), torch.float32)
    u.call('mmlf_dpp_head', u.ptr(scores), u.ptr(bt), u.ptr(bnn), S, Bq, HW, _p(oh.t), _p(po.t), _p(mm.t), _p(lv.t), u.stream())
    [g.check('dpp_head') for g in (oh, po, mm, lv)]
    post = Guarded((Bq, S, 7, 9), torch.float32)
    u.call('mmlf_upr_posterior', _p(mm.t), _p(lv.t), u.ptr(bnn), S, Bq, HW, _p(post.t), u.stream())
    post.check('upr_posterior')
    mask = (torch.rand((Bq, 7, 9), device='cuda') > 0.3).to(torch.int32)
    sm = ops.loss_prepass(mask)
    gsc = Guarded((Bq, S, 7, 9), torch.float32)
    ls = Guarded((1,), torch.float64, fill=torch.zeros(1, dtype=torch.float64, device='cuda'))
    u.call('mmlf_loss_cross_entropy', u.ptr(scores), _p(None), _p(mm.t), u.ptr(bt), 7.0 / 216, S, u.ptr(mask), u.ptr(sm), Bq, HW,
           _p(ls.t), _p(gsc.t), u.stream())
    gsc.check('loss_ce gradient'), ls.check('loss_ce sum')
    K = 7
    means, lvs = torch.randn((K, 1, 7, 9), device='cuda'), torch.randn((K, 1, 7, 9), device='cuda') * 0.3
    em, el, ep = Guarded((1, 7, 9), torch.float32), Guarded((1, 7, 9), torch.float32), Guarded((1, K, 7, 9), torch.float32)
    u.call('mmlf_ese_reduce', u.ptr(means), u.ptr(lvs), u.ptr(ops.numpy_bins(-3.5, 3.5, K, 'cuda')), K, 1, HW, _p(em.t),
           _p(el.t), _p(ep.t), u.stream())
    [g.check('ese_reduce') for g in (em, el, ep)]
    n = 4 * 1000 + 3                                             # not a multiple of the 4-wide vector path
    bufs = [Guarded((n,), torch.float32, fill=torch.rand(n, device='cuda')) for _ in range(3)]
    grad = torch.randn(n, device='cuda')
    u.call('mmlf_adam_step', _p(bufs[0].t), u.ptr(grad), _p(bufs[1].t), _p(bufs[2].t), n, 1e-3, 0.9, 0.999, 1e-8, 1, u.stream())
    [g.check('adam') for g in bufs]


def test_generic_layer_kernels_stay_inside_their_buffers():
    u = _u()
    B, H, W, cin, cout, k, pad = 2, 7, 9, 5, 67, 3, 1
    x = torch.randn((B * H * W, cin), device='cuda')
    wg = torch.randn((k * k * cin, cout), device='cuda')
    y = Guarded((B * H * W, cout), torch.float32)
    u.call('mmlf_g_conv', u.ptr(x), cin, u.ptr(wg), _p(None), B, H, W, cin, cout, k, pad, 1, _p(y.t), cout, u.stream())
    y.check('g_conv')
    dw = Guarded((cout, cin, k, k), torch.float32, fill=torch.zeros((cout, cin, k, k), device='cuda'))
    gy = torch.randn((B * H * W, cout), device='cuda')
    u.call('mmlf_g_conv_wgrad', u.ptr(x), cin, u.ptr(gy), cout, B, H, W, cin, cout, k, pad, 2, 0, _p(dw.t), u.stream())
    dw.check('g_conv_wgrad')
    pool, idx = Guarded((B * (H // 2) * (W // 2), cin), torch.float32), Guarded((B * (H // 2) * (W // 2) * cin,), torch.uint8)
    u.call('mmlf_g_maxpool2', u.ptr(x), B, H, W, cin, _p(pool.t), _p(idx.t), u.stream())
    pool.check('g_maxpool2'), idx.check('g_maxpool2 idx')
    gx = Guarded((B * H * W, cin), torch.float32)
    u.call('mmlf_g_maxpool2_bwd', _p(pool.t), _p(idx.t), B, H, W, cin, _p(gx.t), u.stream())
    gx.check('g_maxpool2_bwd')
    y4 = torch.randn((B * H * W, 4 * 3), device='cuda')
    up = Guarded((B * 2 * H * 2 * W, 3 + 2), torch.float32)
    u.call('mmlf_g_depth_to_space', u.ptr(y4), _p(up.t), 5, 0, B, H, W, 3, 0, u.stream())
    up.check('g_depth_to_space')
    dst = Guarded((B * H * W, 8), torch.float32)
    u.call('mmlf_g_copy_window', u.ptr(x), H, W, cin, 1, 2, 3, _p(dst.t), H, W, 8, 4, 1, 0, B, 4, 5, 3, 0, u.stream())
    dst.check('g_copy_window')
    nchw = Guarded((B, cin, H, W), torch.float32)
    u.call('mmlf_g_layout', _p(nchw.t), u.ptr(x), cin, B, cin, H, W, 0, u.stream())
    nchw.check('g_layout')
    assert torch.equal(nchw.t, x.view(B, H, W, cin).permute(0, 3, 1, 2))
