"""Kernel-level parity tests (-m gpu): every C-ABI entry point against the numpy oracle on seeded inputs."""
import numpy as np
import pytest
import torch

import oracle
from oracle import losses as olosses
from oracle.net import bf16_round, conv2x2, conv2x2_bwd, np_linspace_f32, torch_linspace_f32

pytestmark = pytest.mark.gpu


def _u():
    import _gpu_util
    return _gpu_util


# ------------------------------------------------------------------------------------------------ light field
def test_lf_extract_bit_exact():
    from mmlf_b200.data import hci4d
    rng = np.random.RandomState(0)
    views = rng.randint(0, 256, (81, 24, 32, 3), dtype=np.uint8)
    got = hci4d.extract_stacks(torch.from_numpy(views).cuda())
    want = oracle.extract_stacks(views)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)


@pytest.mark.parametrize('shape', [(1, 9, 16, 16), (2, 9, 12, 20), (1, 9, 33, 18), (1, 9, 40, 512), (2, 9, 96, 96),
                                   (1, 5, 7, 1024), (1, 3, 700, 18)])
def test_lf_shift_bit_exact(shape, golden):
    from mmlf_b200 import ops
    B, n, H, W = shape
    rng = np.random.RandomState(1)
    stacks = [rng.uniform(0, 1, (B, n, 3, H, W)).astype(np.float32) for _ in range(4)]
    if H != W:   # the diagonal stacks assume square images in the reference; h / v are still checked
        pass
    g = golden('shift.npz')
    for disp in list(g['disps']) + [0.37, -2.75]:
        want = oracle.shift(tuple(stacks), float(disp))
        got = ops.lf_shift(*[torch.from_numpy(s).cuda() for s in stacks], float(disp))
        for k in range(4):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (disp, k)


def test_lf_shift_golden(golden):
    """Directly against the reference's own Shift outputs, through the drop-in transform (in-place semantics)."""
    from mmlf_b200.data.hci4d import Shift
    g = golden('shift.npz')
    for j, disp in enumerate(g['disps']):
        data = [torch.from_numpy(g[f'in{k}'].copy()).cuda() for k in range(4)]
        data += [torch.zeros(1), torch.from_numpy(g['gt'].copy()).cuda(), torch.from_numpy(g['mpi'].copy()).cuda()]
        res = Shift(float(disp))(tuple(data))
        for k in range(4):
            assert np.array_equal(res[k].cpu().numpy(), g[f'out{j}_{k}']), (disp, k)
            assert res[k] is data[k]
        assert np.array_equal(res[5].cpu().numpy(), g[f'gt{j}'])
        assert np.array_equal(res[6].cpu().numpy(), g[f'mpi{j}'])


def test_batched_weight_packing_equals_the_per_layer_calls():
    """mmlf_pack_conv_weights_batch (one launch for all layers of a step) == mmlf_pack_conv_weight[_split] per layer."""
    u = _u()
    from mmlf_b200.engine import PackJob
    rng = np.random.RandomState(2)
    cases = [(70, 27, 1, 0, 0, u.FP16), (70, 70, 2, 0, 0, u.FP16), (280, 280, 0, 0, 0, u.BF16), (280, 280, 0, 1, 0, u.BF16),
             (108, 280, 0, 0, 0, u.FP16), (2, 280, 0, 1, 0, u.BF16), (70, 70, 0, 0, 1, u.FP16)]
    jobs, refs, outs, keep = [], [], [], []
    for cout, cin, spatial, dgrad, split, dt in cases:
        w = rng.normal(0, 0.1, (cout, cin, 2, 2)).astype(np.float32)
        b = rng.normal(0, 0.1, cout).astype(np.float32)
        wd, bd = torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()
        n_pad, cin_pad = (u.pad16(cout), u.pad16(cin)) if not dgrad else (u.pad16(cin), u.pad16(cout))
        kc = (cin_pad + 63) // 64
        out = torch.zeros((n_pad, 4 * (3 if split else 1) * kc * 64), dtype=torch.int16, device='cuda')
        bias_pad = torch.full((n_pad,), 7.0, device='cuda')
        if split:
            ref = torch.zeros_like(out)
            u.call('mmlf_pack_conv_weight_split', u.ptr(wd), cout, cin, spatial, 1, cin, u.pad16(cin), u.ptr(ref), n_pad, cin_pad,
                   64.0, u.stream())
        else:
            ref = u.pack_weight(w, spatial=spatial, dgrad=dgrad, n_pad=n_pad, cin_pad=cin_pad, dt=dt).view(torch.int16)
        jobs.append(PackJob(wd.data_ptr(), out.data_ptr(), 0 if dgrad else bd.data_ptr(), 0 if dgrad else bias_pad.data_ptr(),
                            cout, cin, spatial, dgrad, 1, cin, u.pad16(cin), n_pad, cin_pad, dt, split, 64.0 if split else 1.0))
        refs.append((ref, b, dgrad, cout))
        outs.append((out, bias_pad))
        keep += [wd, bd]
    arr = (PackJob * len(jobs))(*jobs)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    max_elems = max(o.numel() for o, _ in outs)
    u.call('mmlf_pack_conv_weights_batch', u.ptr(table), len(jobs), max_elems, u.stream())
    torch.cuda.synchronize()
    for (ref, b, dgrad, cout), (out, bias_pad) in zip(refs, outs):
        assert torch.equal(out, ref.reshape(out.shape))
        if not dgrad:
            bp = bias_pad.cpu().numpy()
            assert np.array_equal(bp[:cout], b) and not bp[cout:].any()


def test_texture_mask(golden):
    from mmlf_b200 import ops
    from mmlf_b200.data import hci4d
    g = golden('texture_mask.npz')
    for tag in ('a', 'b'):
        c, ws, thr = g[f'{tag}/center'], int(g[f'{tag}/wsize']), float(g[f'{tag}/threshold'])
        mask, mae = ops.texture_mask(torch.from_numpy(c).cuda(), ws, thr, want_mae=True)
        np.testing.assert_allclose(mae.cpu().numpy(), g[f'{tag}/mae'], rtol=3e-6, atol=1e-8)
        sure = np.abs(g[f'{tag}/mae'] - np.float32(thr)) > 1e-6
        assert np.array_equal(mask.cpu().numpy()[sure], g[f'{tag}/mask'][sure])
        assert np.array_equal(hci4d.create_mask_texture(torch.from_numpy(c).cuda(), ws, thr).cpu().numpy(), mask.cpu().numpy())
    # full-size image, window 23 (the load_scene call, hci4d.py:241): oracle on a crop
    rng = np.random.RandomState(3)
    big = rng.uniform(0, 1, (1, 3, 512, 512)).astype(np.float32)
    big[..., :200] = 0.5 + 0.01 * big[..., :200]
    mask, mae = ops.texture_mask(torch.from_numpy(big).cuda(), 23, 0.02, want_mae=True)
    want = oracle.texture_mae(big[..., 100:180, 150:260], 23)
    np.testing.assert_allclose(mae.cpu().numpy()[0, 111:169, 161:249], want[0, 11:-11, 11:-11], rtol=3e-6)
    assert not mask[0, :11].any() and not mask[0, :, -11:].any() and mask[0, 11:-11, 300:-11].all()


def test_validation_metrics_against_the_reference(golden):
    from mmlf_b200.validate import metrics as M
    g = golden('metrics.npz')
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    lap = M.laplace_to_discrete(108, -3.5, 3.5, T(g['means'][0]), T(g['logvars'][0]))
    np.testing.assert_allclose(lap.cpu().numpy(), g['laplace'], rtol=1e-5, atol=1e-9)   # float32 exp(logvar): 1 ulp
    lmm = M.lmm_to_discrete(108, -3.5, 3.5, T(g['means']), T(g['logvars']))
    np.testing.assert_allclose(lmm.cpu().numpy(), g['lmm'], rtol=1e-5, atol=1e-9)
    assert abs(lmm.sum(1).mean().item() - g['lmm'].sum(1).mean()) < 1e-7
    assert np.array_equal(M.mean_to_discrete(108, -3.5, 3.5, T(g['means'][0])).cpu().numpy(), g['mean_disc'])
    mask = M.multimodal_mask(T(g['mpi']))
    assert np.array_equal(mask.cpu().numpy(), g['mm_mask'])
    d, gt = T(g['lmm'].copy()), T(g['dist_gt'].copy())            # the reference's own distribution: exact comparison
    vals = [M.kl_divergence(d, gt), M.kl_divergence(d, gt, mask), M.kl_divergence(d, gt, 1.0 - mask)]
    np.testing.assert_allclose(vals, g['kld'], rtol=1e-12)
    np.testing.assert_allclose(d.cpu().numpy(), g['kld_dist_after'], rtol=1e-12)
    np.testing.assert_allclose(gt.cpu().numpy(), g['kld_gt_after'], rtol=1e-12)
    w, p = T(g['dist_gt'].copy()), T(g['lmm'].copy())
    np.testing.assert_allclose(M.nll_discrete(w, p), float(g['nll']), rtol=1e-12)
    # full size: 70 members, one 512 x 512 light field, mass conservation (the CDF differences telescope)
    gen = torch.Generator(device='cuda').manual_seed(0)
    means = torch.rand((70, 1, 512, 512), device='cuda', generator=gen) * 4 - 2
    logvars = torch.randn((70, 1, 512, 512), device='cuda', generator=gen) * 0.5 - 1
    big = M.lmm_to_discrete(108, -3.5, 3.5, means, logvars)
    assert big.min().item() >= 0 and big.sum(1).max().item() <= 1 + 1e-12 and big.sum(1).min().item() > 0.5


def test_augmentation_chain_against_the_reference(golden):
    """GPU augmentation chain (one gather kernel + Contrast) against the reference's own Compose output: bit exact for
    every tensor when Contrast is given NumPy's float32 mean; with the on-device float64 mean the views agree to 1 ulp."""
    import random
    from oracle import augment as A
    from mmlf_b200.data.augment import GpuAugmenter, draw_params
    g = golden('augment.npz')
    sample = [g[k] for k in ('h', 'v', 'i', 'd', 'center', 'gt', 'mpi', 'mask')] + [np.atleast_1d(3)]
    aug = GpuAugmenter([sample, sample])
    seeds = [int(s) for s in g['seeds']]
    params = [draw_params(random.Random(s), 64, 64, int(g['ps']), int(g['max_factor'])) for s in seeds]
    means = [A.augment(sample, p)[1] for p in params]
    ids = [i % 2 for i in range(len(seeds))]
    out = aug(ids, params, mean_override=means)
    names = ('h', 'v', 'i', 'd', 'center', 'gt', 'mpi', 'mask')
    for j, s in enumerate(seeds):
        for k, name in enumerate(names):
            ref = g[f'{s}/{name}']
            got = out[k][j].cpu().numpy()
            if name == 'mpi':
                ref = ref.astype(np.float32)                    # train/cli.py casts with .float() after the transforms
            assert np.array_equal(got, ref.astype(got.dtype)), (s, name, np.abs(got - ref).max())
    # on-device mean: float64 sum instead of NumPy's pairwise float32 sum
    out2 = aug(ids, params)
    dm = np.abs(aug.last_means.cpu().numpy() - np.array(means, np.float64)) / np.array(means, np.float64)
    assert dm.max() < 1e-6
    for k in range(5):
        a, b = out2[k].cpu().numpy(), out[k].cpu().numpy()
        assert np.abs(a - b).max() <= 2.4e-7 * max(1.0, np.abs(b).max())
    for k in (5, 6, 7):
        assert torch.equal(out2[k], out[k])


def test_augmentation_chain_multi_tile_against_the_oracle():
    """The same chain at a patch size that spans several 16 x 16 tiles with a ragged edge (ps = 40), every rotation and
    down-sampling factor, against the numpy oracle (pinned to the reference by the golden of the test above): bit exact
    with the oracle's float32 Contrast mean."""
    import random
    from oracle import augment as A
    from mmlf_b200.data.augment import GpuAugmenter, draw_params
    rs = np.random.RandomState(5)
    H = W = 184                                                  # (40 + 16) * 3 = 168 fits with max_factor 3
    sample = [rs.uniform(0, 1, (9, 3, H, W)).astype(np.float32) for _ in range(4)]
    sample += [sample[1][4].copy(), rs.uniform(-2, 2, (H, W)).astype(np.float32),
               rs.uniform(-2, 2, (2, 5, H, W)).astype(np.float32).astype(np.float64),      # float64 array of float32 values
               (rs.uniform(0, 1, (H, W)) > 0.3).astype(np.int32), np.atleast_1d(7)]
    aug = GpuAugmenter([sample])
    params, seen = [], set()
    seed = 0
    while len(seen) < 12 and seed < 400:                         # every (rotation, factor) pair once
        p = draw_params(random.Random(seed), H, W, 40, 3)
        seed += 1
        if (p['r'], p['f']) not in seen:
            seen.add((p['r'], p['f']))
            params.append(p)
    assert len(seen) == 12
    refs = [A.augment(sample, p) for p in params]
    out = aug([0] * len(params), params, mean_override=[m for _, m in refs])
    for j, (ref, _) in enumerate(refs):
        for k, name in enumerate(('h', 'v', 'i', 'd', 'center', 'gt', 'mpi', 'mask')):
            want = ref[k].astype(np.float32) if name == 'mpi' else ref[k]
            got = out[k][j].cpu().numpy()
            assert np.array_equal(got, want.astype(got.dtype)), (params[j]['r'], params[j]['f'], name)


def test_pack_views_and_shift_pack():
    u = _u()
    rng = np.random.RandomState(2)
    for B, n, H, W, dt in ((2, 9, 10, 14, u.BF16), (2, 9, 10, 14, u.FP16), (1, 9, 5, 300, u.FP16)):
        v = rng.uniform(0, 1, (B, n, 3, H, W)).astype(np.float32)
        out = torch.full((B * (H + 1) * (W + 1), 32), float('nan'), dtype=u.TDT[dt], device='cuda')
        vd = torch.from_numpy(v).cuda()                     # named: a temporary would be freed before the launch
        u.call('mmlf_pack_views', u.ptr(vd), B, n * 3, H, W, u.ptr(out), 32, dt, u.stream())
        got = out.float().cpu().numpy().reshape(B, H + 1, W + 1, 32)
        want = np.zeros_like(got)
        want[:, 1:, 1:, :27] = u.ROUND[dt](v.reshape(B, 27, H, W).transpose(0, 2, 3, 1))
        assert np.array_equal(got, want)
    # fused shift + pack == pack(shift(.)) for every stack
    B, n = 2, 9
    for Hs, Ws in ((12, 12), (9, 150), (8, 136)):
        stacks = [rng.uniform(0, 1, (B, n, 3, Hs, Ws)).astype(np.float32) for _ in range(4)]
        for disp in (2.5, -1.3, 0.0, 7.25):
            sh = oracle.shift(tuple(stacks), disp)
            for k in range(4):
                o = torch.full((B * (Hs + 1) * (Ws + 1), 32), float('nan'), dtype=torch.float16, device='cuda')
                sd = torch.from_numpy(stacks[k]).cuda()
                u.call('mmlf_shift_pack', u.ptr(sd), k, B, n, Hs, Ws, float(disp),
                       u.ptr(o), 32, u.FP16, u.stream())
                got = o.float().cpu().numpy().reshape(B, Hs + 1, Ws + 1, 32)
                want = np.zeros_like(got)
                want[:, 1:, 1:, :27] = u.ROUND[u.FP16](sh[k].reshape(B, 27, Hs, Ws).transpose(0, 2, 3, 1))
                assert np.array_equal(got, want), (disp, k, Hs, Ws)


# ------------------------------------------------------------------------------------------------ convolution
def _conv_case(B, H, W, cin, cout, ctype, seed, simt, relu=True, mode=0, with_bn=False, groups=1, group_real=None,
               group_pad=None, dt=1):
    u = _u()
    rnd = u.ROUND[dt]
    rng = np.random.RandomState(seed)
    Hp, Wp = H + 1, W + 1
    cin_pad = groups * group_pad if groups > 1 else u.pad16(cin)
    n_pad = u.pad16(cout)
    w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    b = rng.normal(0, 0.1, cout).astype(np.float32)
    if ctype == 0:
        x = rng.normal(0, 1, (B, H, W, cin)).astype(np.float32)
    else:
        x = rng.normal(0, 1, (B, Hp, Wp, cin)).astype(np.float32)
    xq, wq = rnd(x), rnd(w)
    want = conv2x2(xq, wq, b, 1 if ctype == 0 else 0)
    scale = shift = None
    if with_bn:
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        shift = rng.normal(0, 0.2, cout).astype(np.float32)
        want = want * scale + shift
    if relu:
        want = np.maximum(want, 0)
    # channel layout on the device (groups of group_real real channels at pitch group_pad)
    if groups > 1:
        xl = np.zeros(x.shape[:3] + (cin_pad,), np.float32)
        for g in range(groups):
            xl[..., g * group_pad:g * group_pad + group_real] = xq[..., g * group_real:(g + 1) * group_real]
    else:
        xl = xq
    xs = u.to_slots(xl, cin_pad, ctype == 1, Hp, Wp, dt)
    wp = u.pack_weight(w, groups=groups, group_real=group_real, group_pad=group_pad, dt=dt)
    bias = torch.zeros(n_pad, device='cuda')
    bias[:cout] = torch.from_numpy(b)
    kw = dict(bias=bias, relu=relu, simt=simt, out_mode=mode, ab=dt, out_dt=dt)
    if with_bn:
        sc = torch.zeros(n_pad, device='cuda')
        sh = torch.zeros(n_pad, device='cuda')
        sc[:cout] = torch.from_numpy(scale)
        sh[:cout] = torch.from_numpy(shift)
        kw.update(scale=sc, shift=sh)
    if mode == 2:
        kw['n_real'] = cout
    out = u.run_conv(xs, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, **kw)
    if mode == 2:
        got = out.cpu().numpy().transpose(0, 2, 3, 1)
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)
        return
    full = out.float().cpu().numpy().reshape(B, Hp, Wp, -1)
    assert np.isfinite(full).all(), 'kernel left unwritten (NaN) output slots'
    got = full[..., :cout] if ctype == 0 else full[:, 1:, 1:, :cout]
    if ctype == 1:
        assert not full[:, 0].any() and not full[:, :, 0].any(), 'halo slots must be zero'
    assert not full[..., cout:].any() or not relu, 'padding channels must stay zero'
    if mode == 0:
        u.assert_close_bf16(got, want, f'conv type {ctype} {cin}->{cout}', ulps=1.01, atol=2e-3 if dt == 0 else 3e-4,
                            dt=dt)
    else:
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('case', [(2, 12, 12, 27, 70, 0), (1, 20, 24, 280, 280, 1), (2, 9, 11, 70, 70, 1),
                                  (1, 16, 16, 280, 108, 0)])
def test_conv_split_precision(case):
    """Split-precision conv (hi*hi + hi*lo + lo*hi on the fp16 tensor cores, fp32 accumulate): fp32-class agreement with
    a float64 convolution of the fp32 operands, and the hi / lo outputs re-assemble the fp32 result."""
    u = _u()
    B, H, W, cin, cout, ctype = case
    rng = np.random.RandomState(11)
    Hp, Wp = H + 1, W + 1
    cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
    w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    b = rng.normal(0, 0.1, cout).astype(np.float32)
    x = rng.normal(0, 1, (B, H, W, cin) if ctype == 0 else (B, Hp, Wp, cin)).astype(np.float32)
    want = np.maximum(conv2x2(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), 1 if ctype == 0 else 0), 0)
    hi = u.ROUND[u.FP16](x)
    lo = u.ROUND[u.FP16](x - hi)
    xs = torch.cat([u.to_slots(hi, cin_pad, ctype == 1, Hp, Wp, u.FP16), u.to_slots(lo, cin_pad, ctype == 1, Hp, Wp, u.FP16)], 1)
    kc = (cin_pad + 63) // 64
    wp = torch.empty((n_pad, 4 * 3 * kc * 64), dtype=torch.float16, device='cuda')
    wd = torch.from_numpy(w).cuda()
    u.call('mmlf_pack_conv_weight_split', u.ptr(wd), cout, cin, 0, 1, cin, cin_pad, u.ptr(wp), n_pad,
           cin_pad, 64.0, u.stream())
    bias = torch.zeros(n_pad, device='cuda')
    bias[:cout] = torch.from_numpy(b)
    unscale = torch.full((n_pad,), 1.0 / 64.0, device='cuda')
    n_slots = B * Hp * Wp
    out = torch.full((n_slots, 2 * n_pad), float('nan'), dtype=torch.float16, device='cuda')
    a = u.ConvArgs()
    a.in_, a.ld_in, a.cin_pad, a.wpack, a.n_pad = xs.data_ptr(), 2 * cin_pad, cin_pad, wp.data_ptr(), n_pad
    a.B, a.H, a.W, a.type = B, H, W, ctype
    a.scale, a.shift, a.relu = unscale.data_ptr(), bias.data_ptr(), 1
    a.out, a.ld_out, a.out_mode = out.data_ptr(), 2 * n_pad, 0
    a.ab_dtype, a.out_dtype, a.out2_dtype = u.FP16, u.FP16, u.FP16
    a.split_in, a.split_out = cin_pad, n_pad
    u.call('mmlf_conv2x2', u.C.byref(a), u.stream())
    torch.cuda.synchronize()
    full = out.float().cpu().numpy().reshape(B, Hp, Wp, 2, n_pad)
    assert np.isfinite(full).all()
    got = full[..., 0, :] + full[..., 1, :]                  # hi + lo
    got = got[..., :cout] if ctype == 0 else got[:, 1:, 1:, :cout]
    err = np.abs(got - want).max() / np.abs(want).max()
    # measured 2e-7 (K = 108) ... 5.6e-6 (K = 1120): what is left is the tensor cores' fp32 accumulation, which truncates
    # each of the K / 16 additions instead of rounding to nearest; the plain fp16 path is at ~5e-4 on the same data
    assert err < 1e-5, err
    if ctype == 1:
        assert not full[:, 0].any() and not full[:, :, 0].any(), 'halo slots must be zero'


@pytest.mark.parametrize('case', [(2, 12, 12, 280, 280, 1), (3, 9, 11, 70, 70, 1), (1, 20, 24, 280, 280, 0)])
def test_conv_fused_bn_backward_statistics(case):
    """Data-gradient launch with fused BatchNorm-backward statistics == mmlf_bn_bwd_reduce on the stored gradient."""
    u = _u()
    B, H, W, cin, cout, ctype = case
    rng = np.random.RandomState(5)
    Hp, Wp = H + 1, W + 1
    cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
    n_slots = B * Hp * Wp
    w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    x = rng.normal(0, 1, (B, H, W, cin) if ctype == 0 else (B, Hp, Wp, cin)).astype(np.float32)
    xs = u.to_slots(x, cin_pad, ctype == 1, Hp, Wp, u.BF16)
    wp = u.pack_weight(w, dt=u.BF16)
    z = (torch.randn((n_slots, n_pad), device='cuda') * 1.5 + 0.3).to(torch.float16)
    scale = torch.zeros(n_pad, device='cuda')
    shift = torch.zeros(n_pad, device='cuda')
    mean = torch.zeros(n_pad, device='cuda')
    invstd = torch.zeros(n_pad, device='cuda')
    scale[:cout] = torch.from_numpy(rng.uniform(-1.5, 1.5, cout).astype(np.float32))
    shift[:cout] = torch.from_numpy(rng.normal(0, 0.5, cout).astype(np.float32))
    mean[:cout] = torch.from_numpy(rng.normal(0.3, 0.2, cout).astype(np.float32))
    invstd[:cout] = torch.from_numpy(rng.uniform(0.5, 2.0, cout).astype(np.float32))
    fused = torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda')
    out = torch.full((n_slots, n_pad), float('nan'), dtype=torch.bfloat16, device='cuda')
    a = u.ConvArgs()
    a.in_, a.ld_in, a.cin_pad, a.wpack, a.n_pad = xs.data_ptr(), cin_pad, cin_pad, wp.data_ptr(), n_pad
    a.B, a.H, a.W, a.type = B, H, W, ctype
    a.out, a.ld_out, a.out_mode = out.data_ptr(), n_pad, 0
    a.col_sums = fused.data_ptr()
    a.ab_dtype, a.out_dtype, a.out2_dtype = u.BF16, u.BF16, u.BF16
    a.bn_z, a.ld_z, a.bn_z_dtype = z.data_ptr(), n_pad, u.FP16
    a.bn_scale, a.bn_shift, a.bn_mean = scale.data_ptr(), shift.data_ptr(), mean.data_ptr()
    u.call('mmlf_conv2x2', u.C.byref(a), u.stream())
    # the plain data-gradient variant must produce the same gradient
    plain = u.run_conv(xs, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, ab=u.BF16, out_dt=u.BF16)
    assert torch.equal(out, plain)
    want = torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda')
    u.call('mmlf_bn_bwd_reduce', u.ptr(out), n_pad, u.ptr(z), n_pad, u.ptr(scale), u.ptr(shift), u.ptr(mean), u.ptr(invstd),
           n_pad, B, H, W, u.BF16, u.FP16, u.ptr(want), u.stream())
    torch.cuda.synchronize()
    f, wnt = fused.cpu().numpy().reshape(2, n_pad), want.cpu().numpy().reshape(2, n_pad)
    np.testing.assert_allclose(f[0], wnt[0], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(f[1] * invstd.cpu().numpy(), wnt[1], rtol=2e-4, atol=4e-3)


CONV_CASES = [
    # B, H, W, cin, cout, type
    (2, 12, 12, 27, 70, 0), (2, 12, 12, 70, 70, 1), (1, 20, 24, 280, 280, 0), (1, 20, 24, 280, 280, 1),
    (1, 16, 16, 280, 2, 0), (1, 16, 16, 280, 108, 0), (1, 16, 16, 108, 108, 1), (2, 9, 11, 140, 140, 0),
]


@pytest.mark.parametrize('dt', [0, 1])
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_simt_vs_oracle(case, dt):
    _conv_case(*case, seed=3, simt=True, dt=dt)


@pytest.mark.parametrize('dt', [0, 1])
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_tc_vs_oracle(case, dt):
    _conv_case(*case, seed=3, simt=False, dt=dt)


def test_conv_tc_epilogues():
    _conv_case(1, 16, 16, 280, 280, 1, seed=4, simt=False, with_bn=True)
    _conv_case(1, 16, 16, 280, 280, 1, seed=5, simt=False, relu=False)
    _conv_case(1, 16, 16, 280, 16, 0, seed=6, simt=False, mode=1)
    _conv_case(1, 16, 16, 108, 108, 1, seed=7, simt=False, relu=False, mode=2)
    _conv_case(1, 16, 16, 280, 280, 0, seed=8, simt=False, groups=4, group_real=70, group_pad=80)


def test_conv_tc_many_tiles():
    """More tiles than SMs and more k-chunks than pipeline stages: exercises barrier phase wrap-around."""
    _conv_case(4, 80, 80, 280, 280, 0, seed=9, simt=False)
    _conv_case(4, 80, 80, 280, 280, 1, seed=10, simt=False)


def test_conv_dgrad_and_gate():
    """Data gradient = the other conv type with dgrad-packed weights; ReLU gate fused in the epilogue."""
    u = _u()
    rng = np.random.RandomState(11)
    B, H, W, cin, cout = 2, 10, 12, 70, 70
    Hp, Wp = H + 1, W + 1
    w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    wq = bf16_round(w)
    for ctype in (0, 1):
        if ctype == 0:
            x = rng.normal(0, 1, (B, H, W, cin)).astype(np.float32)
            gout = bf16_round(rng.normal(0, 1, (B, Hp, Wp, cout)).astype(np.float32))
        else:
            x = rng.normal(0, 1, (B, Hp, Wp, cin)).astype(np.float32)
            gout = bf16_round(rng.normal(0, 1, (B, H, W, cout)).astype(np.float32))
        gx, _, _ = conv2x2_bwd(bf16_round(x), wq, gout, 1 if ctype == 0 else 0)
        gate = (rng.uniform(size=gx.shape) > 0.4).astype(np.float32)
        want = gx * gate
        wd = u.pack_weight(w, dgrad=1, n_pad=u.pad16(cin), cin_pad=u.pad16(cout))
        gs = u.to_slots(gout, u.pad16(cout), ctype == 0, Hp, Wp)
        gate_s = u.to_slots(gate, u.pad16(cin), ctype == 1, Hp, Wp, u.FP16)
        bits = u.pack_bits(gate_s.float().cpu().numpy() > 0, ld_bits=4)
        for simt in (True, False):
            out = u.run_conv(gs, u.pad16(cout), u.pad16(cout), wd, u.pad16(cin), B, H, W, 1 - ctype, gate_bits=bits,
                             ld_bits=4, simt=simt)
            got = u.from_slots(out, B, Hp, Wp, cin, ctype == 1)
            u.assert_close_bf16(got, want, f'dgrad of type {ctype} (simt={simt})', ulps=1.01, atol=2e-3)
        got = u.from_slots(out, B, Hp, Wp, cin, ctype == 1)
        u.assert_close_bf16(got, want, f'dgrad of type {ctype}', ulps=1.01, atol=2e-3)


@pytest.mark.parametrize('case', [(3, 40, 40, 280, 280, 0, 1), (3, 40, 40, 280, 280, 1, 1), (2, 24, 20, 70, 70, 1, 1),
                                  (2, 24, 20, 27, 70, 0, 0), (1, 30, 30, 280, 108, 0, 1), (2, 17, 13, 320, 320, 1, 0)])
def test_conv_fused_epilogue_outputs(case):
    """Second output copy in the other 16-bit format, ReLU sign bits and per-channel sum / sum of squares all come
    out of the same epilogue and must describe exactly the stored primary output."""
    u = _u()
    B, H, W, cin, cout, ctype, dt = case
    rng = np.random.RandomState(21)
    Hp, Wp = H + 1, W + 1
    n_pad, cin_pad = u.pad16(cout), u.pad16(cin)
    w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    b = rng.normal(0, 0.3, cout).astype(np.float32)
    shp = (B, H, W, cin) if ctype == 0 else (B, Hp, Wp, cin)
    xs = u.to_slots(rng.normal(0, 1, shp).astype(np.float32), cin_pad, ctype == 1, Hp, Wp, dt)
    wp = u.pack_weight(w, dt=dt)
    bias = torch.zeros(n_pad, device='cuda')
    bias[:cout] = torch.from_numpy(b)
    ld_bits = (n_pad + 31) // 32
    n_slots = B * Hp * Wp
    ref = u.run_conv(xs, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, bias=bias, relu=True, ab=dt, out_dt=dt)
    rbits = torch.full((n_slots, ld_bits), -1, dtype=torch.int32, device='cuda')
    sums = torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda')
    out, out2 = u.run_conv(xs, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, bias=bias, relu=True, ab=dt, out_dt=dt,
                           out2_dt=1 - dt, relu_bits=rbits, ld_bits=ld_bits, col_sums=sums)
    assert torch.equal(out, ref), 'primary output must not depend on the optional epilogue outputs'
    o = out.float()
    want2 = o.to(u.TDT[1 - dt])
    if 1 - dt == u.FP16:      # the kernel converts the fp32 value, not the rounded primary: allow one fp16 ulp
        d = (out2.float() - o).abs()
        assert bool((d <= o.abs() * 2.0 ** -7 + 1e-6).all()), 'second copy differs from the primary by more than a bf16 ulp'
    else:
        d = (out2.float() - o).abs()
        assert bool((d <= o.abs() * 2.0 ** -7 + 1e-6).all()), 'bf16 copy differs from the fp16 primary by more than a bf16 ulp'
    assert want2.shape == out2.shape
    got_bits = u.unpack_bits(rbits, n_pad)
    assert np.array_equal(got_bits, (o > 0).cpu().numpy()), 'ReLU bits do not match the stored output'
    assert not u.unpack_bits(rbits, ld_bits * 32)[:, n_pad:].any()
    od = o.double()
    s1, s2 = od.sum(0).cpu().numpy(), (od * od).sum(0).cpu().numpy()
    gs = sums.cpu().numpy()
    np.testing.assert_allclose(gs[:n_pad], s1, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(gs[n_pad:], s2, rtol=2e-5, atol=1e-3)


@pytest.mark.parametrize('case', [(2, 12, 12, 27, 70, 0), (2, 12, 12, 70, 70, 1), (2, 20, 20, 280, 280, 0),
                                  (2, 20, 20, 280, 280, 1), (1, 16, 16, 280, 2, 0), (1, 16, 16, 108, 108, 1),
                                  (6, 40, 40, 280, 280, 1)])
@pytest.mark.parametrize('act_dt', [0, 1])
def test_conv_wgrad(case, act_dt):
    """act_dt = 1: fp16 activations are converted to bf16 first (mixed-format MMA operands fault on B200)."""
    u = _u()
    rnd = u.ROUND[act_dt]
    B, H, W, cin, cout, ctype = case
    rng = np.random.RandomState(12)
    Hp, Wp = H + 1, W + 1
    cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
    if ctype == 0:
        x = rnd(rng.normal(0, 1, (B, H, W, cin)).astype(np.float32))
        gout = bf16_round(rng.normal(0, 1, (B, Hp, Wp, cout)).astype(np.float32))
    else:
        x = rnd(rng.normal(0, 1, (B, Hp, Wp, cin)).astype(np.float32))
        gout = bf16_round(rng.normal(0, 1, (B, H, W, cout)).astype(np.float32))
    w = np.zeros((cout, cin, 2, 2), np.float32)
    _, gw, gb = conv2x2_bwd(x, w, gout, 1 if ctype == 0 else 0)
    xs = u.to_slots(x, cin_pad, ctype == 1, Hp, Wp, act_dt)
    gs = u.to_slots(gout, n_pad, ctype == 0, Hp, Wp)
    ws_bytes = u._lib.lib().mmlf_conv2x2_wgrad_workspace(n_pad, cin_pad)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device='cuda')
    dw_mixed = None
    if act_dt == u.FP16:
        x = bf16_round(x)
        _, gw, gb = conv2x2_bwd(x, w, gout, 1 if ctype == 0 else 0)
        # mixed formats straight into the kernel: the fp16 activation boxes are converted to bf16 in shared memory
        dw_mixed = torch.full((n_pad, 4, cin_pad), float('nan'), dtype=torch.float32, device='cuda')
        u.call('mmlf_conv2x2_wgrad', u.ptr(gs), n_pad, n_pad, u.ptr(xs), cin_pad, cin_pad, B, H, W, ctype, u.FP16, u.BF16,
               u.ptr(ws), u.ptr(dw_mixed), u.stream())
        xb = torch.full_like(xs, float('nan'), dtype=torch.bfloat16)
        u.call('mmlf_convert16', u.ptr(xs), cin_pad, u.FP16, u.ptr(xb), cin_pad, u.BF16, cin_pad, B * Hp * Wp, u.stream())
        xs = xb
    dwp = torch.full((n_pad, 4, cin_pad), float('nan'), dtype=torch.float32, device='cuda')
    u.call('mmlf_conv2x2_wgrad', u.ptr(gs), n_pad, n_pad, u.ptr(xs), cin_pad, cin_pad, B, H, W, ctype, u.BF16, u.BF16,
           u.ptr(ws), u.ptr(dwp), u.stream())
    dw = torch.empty((cout, cin, 2, 2), dtype=torch.float32, device='cuda')
    u.call('mmlf_unpack_conv_wgrad', u.ptr(dwp), n_pad, cin_pad, cout, cin, 0, 1, cin, cin_pad, u.ptr(dw), 0, u.stream())
    torch.cuda.synchronize()
    if dw_mixed is not None:                 # same operand values, same MMAs: bit-identical to convert-then-multiply
        assert torch.equal(dw_mixed, dwp), 'in-kernel fp16 -> bf16 conversion differs from mmlf_convert16 + wgrad'
    scale = np.abs(gw).max()
    assert np.abs(dw.cpu().numpy() - gw).max() <= 2e-4 * scale + 1e-4
    db = torch.zeros(n_pad, device='cuda')
    u.call('mmlf_colsum16', u.ptr(gs), n_pad, n_pad, B * Hp * Wp, u.BF16, u.ptr(db), 0, u.stream())
    np.testing.assert_allclose(db.cpu().numpy()[:cout], gb, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('case', [(2, 12, 12, 70, 70, 1, 1, 1), (2, 12, 12, 27, 70, 0, 2, 1), (1, 20, 20, 280, 280, 0, 0, 4),
                                  (1, 16, 16, 280, 2, 0, 0, 1)])
def test_conv_wgrad_canonical_equals_wgrad_plus_unpack(case):
    """mmlf_conv2x2_wgrad_canonical (K-split reduction straight into the (cout, cin, 2, 2) gradient, per-stream tap
    mapping and channel groups undone) == mmlf_conv2x2_wgrad + mmlf_unpack_conv_wgrad bit for bit, also accumulating."""
    u = _u()
    B, H, W, cin, cout, ctype, spatial, groups = case
    rng = np.random.RandomState(19)
    Hp, Wp = H + 1, W + 1
    group_real = cin // groups
    group_pad = u.pad16(group_real)
    cin_pad, n_pad = groups * group_pad, u.pad16(cout)
    n_slots = B * Hp * Wp
    xs = torch.from_numpy(rng.normal(0, 1, (n_slots, cin_pad)).astype(np.float32)).cuda().to(torch.bfloat16)
    gs = torch.from_numpy(rng.normal(0, 1, (n_slots, n_pad)).astype(np.float32)).cuda().to(torch.bfloat16)
    ws = torch.empty(u._lib.lib().mmlf_conv2x2_wgrad_workspace(n_pad, cin_pad) // 4, dtype=torch.float32, device='cuda')
    dwp = torch.empty((n_pad, 4, cin_pad), dtype=torch.float32, device='cuda')
    base = torch.from_numpy(rng.normal(0, 1, (cout, cin, 2, 2)).astype(np.float32)).cuda()
    for accumulate in (0, 1):
        want, got = base.clone(), base.clone()
        u.call('mmlf_conv2x2_wgrad', u.ptr(gs), n_pad, n_pad, u.ptr(xs), cin_pad, cin_pad, B, H, W, ctype, u.BF16, u.BF16,
               u.ptr(ws), u.ptr(dwp), u.stream())
        u.call('mmlf_unpack_conv_wgrad', u.ptr(dwp), n_pad, cin_pad, cout, cin, spatial, groups, group_real, group_pad,
               u.ptr(want), accumulate, u.stream())
        u.call('mmlf_conv2x2_wgrad_canonical', u.ptr(gs), n_pad, n_pad, u.ptr(xs), cin_pad, cin_pad, B, H, W, ctype, u.BF16,
               u.BF16, u.ptr(ws), cout, cin, spatial, groups, group_real, group_pad, u.ptr(got), accumulate, u.stream())
        torch.cuda.synchronize()
        assert torch.equal(got, want), (case, accumulate)
        assert not torch.equal(got, base)
        # a destination that is only 4-byte aligned (a slice of the flat gradient buffer behind an odd-sized bias)
        buf = torch.full((base.numel() + 8,), 7.0, dtype=torch.float32, device='cuda')
        odd = buf[1:1 + base.numel()].view(base.shape)
        odd.copy_(base)
        u.call('mmlf_conv2x2_wgrad_canonical', u.ptr(gs), n_pad, n_pad, u.ptr(xs), cin_pad, cin_pad, B, H, W, ctype, u.BF16,
               u.BF16, u.ptr(ws), cout, cin, spatial, groups, group_real, group_pad, odd.data_ptr(), accumulate, u.stream())
        torch.cuda.synchronize()
        assert torch.equal(odd, want), (case, accumulate, 'unaligned')
        assert buf[0].item() == 7.0 and bool((buf[1 + base.numel():] == 7.0).all())


def test_weight_pack_variants():
    """The three spatial variants against the reference's permute/flip plumbing (feed_forward.py:236-256)."""
    u = _u()
    rng = np.random.RandomState(13)
    B, H, W, cin, cout = 1, 8, 8, 27, 16
    w = rng.normal(0, 0.3, (cout, cin, 2, 2)).astype(np.float32)
    b = np.zeros(cout, np.float32)
    x = bf16_round(rng.normal(0, 1, (B, H, W, cin)).astype(np.float32))
    wq = bf16_round(w)
    for spatial in (0, 1, 2):
        if spatial == 0:
            want = conv2x2(x, wq, b, 1)
        elif spatial == 1:          # permute(0,1,3,2) -> net -> permute back
            want = conv2x2(x.transpose(0, 2, 1, 3), wq, b, 1).transpose(0, 2, 1, 3)
        else:                       # permute, flip(-1) -> net -> flip(-1), permute
            xi = np.ascontiguousarray(x.transpose(0, 2, 1, 3)[:, :, ::-1])
            want = conv2x2(xi, wq, b, 1)[:, :, ::-1].transpose(0, 2, 1, 3)
        wp = u.pack_weight(w, spatial=spatial)
        xs = u.to_slots(x, 32, False, H + 1, W + 1)
        out = u.run_conv(xs, 32, 32, wp, 16, B, H, W, 0, simt=True, out_mode=1)
        got = out.cpu().numpy().reshape(B, H + 1, W + 1, 16)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4, err_msg=f'spatial {spatial}')


# ------------------------------------------------------------------------------------------------ batch norm
@pytest.mark.parametrize('shape', [(2, 9, 13, 70), (1, 96, 96, 280)])
def test_bn_train_roundtrip(shape):
    """BatchNorm forward (statistics, finalize, apply + ReLU) and backward (reduce, apply) against float64 numpy; the
    second shape is one BASELINE-size patch of the out-net (96 x 96 px, 280 channels)."""
    u = _u()
    rng = np.random.RandomState(14)
    B, H, W, Cr = shape
    Cp = u.pad16(Cr)
    Hp, Wp = H + 1, W + 1
    A, G = u.FP16, u.BF16
    z = u.ROUND[A](rng.normal(0.3, 1.7, (B, H, W, Cr)).astype(np.float32))
    zs = u.to_slots(z, Cp, False, Hp, Wp, A)
    sums = torch.zeros(2 * Cp, dtype=torch.float64, device='cuda')
    u.call('mmlf_bn_stats', u.ptr(zs), Cp, Cp, B, H, W, A, u.ptr(sums), u.stream())
    gamma = rng.uniform(0.5, 1.5, Cr).astype(np.float32)
    beta = rng.normal(0, 0.2, Cr).astype(np.float32)
    rm0 = rng.normal(0, 0.1, Cr).astype(np.float32)
    rv0 = rng.uniform(0.5, 1.5, Cr).astype(np.float32)
    d = {k: u.dev_f32(v) for k, v in dict(gamma=gamma, beta=beta, rm=rm0, rv=rv0).items()}
    nbt = torch.zeros((), dtype=torch.int64, device='cuda')
    scale, shift, smean, sinv = [torch.empty(Cp, device='cuda') for _ in range(4)]
    n = B * H * W
    u.call('mmlf_bn_finalize', u.ptr(sums), Cr, Cp, n, u.ptr(d['gamma']), u.ptr(d['beta']), u.ptr(d['rm']),
           u.ptr(d['rv']), u.ptr(nbt), 0.1, 1e-5, u.ptr(scale), u.ptr(shift), u.ptr(smean), u.ptr(sinv), u.stream())
    y = torch.full((B * Hp * Wp, Cp), float('nan'), dtype=torch.float16, device='cuda')
    y2 = torch.full((B * Hp * Wp, Cp), float('nan'), dtype=torch.bfloat16, device='cuda')
    u.call('mmlf_bn_apply_relu', u.ptr(zs), Cp, u.ptr(scale), u.ptr(shift), Cp, B, H, W, A, u.ptr(y), Cp, u.ptr(y2), Cp, G,
           u.stream())
    torch.cuda.synchronize()
    assert bool(((y2.float() - y.float()).abs() <= y.float().abs() * 2.0 ** -7 + 1e-6).all()), 'bf16 copy of y is off'
    assert bool(((y2 > 0) == (y > 0)).all())
    z2 = z.reshape(-1, Cr).astype(np.float64)
    mean, var = z2.mean(0), z2.var(0)
    np.testing.assert_allclose(smean.cpu().numpy()[:Cr], mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sinv.cpu().numpy()[:Cr], 1 / np.sqrt(var + 1e-5), rtol=1e-5)
    np.testing.assert_allclose(d['rm'].cpu().numpy(), 0.9 * rm0 + 0.1 * mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d['rv'].cpu().numpy(), 0.9 * rv0 + 0.1 * var * n / (n - 1), rtol=1e-5)
    assert int(nbt.item()) == 1
    want = np.maximum((z - mean) / np.sqrt(var + 1e-5) * gamma + beta, 0).astype(np.float32)
    full = y.float().cpu().numpy().reshape(B, Hp, Wp, Cp)
    assert not full[:, 0].any() and not full[:, :, 0].any() and not full[..., Cr:].any()
    u.assert_close_bf16(full[:, 1:, 1:, :Cr], want, 'bn_apply_relu', ulps=1.01, atol=3e-4, dt=A)
    # backward
    gy = bf16_round(rng.normal(0, 1, (B, H, W, Cr)).astype(np.float32))
    gys = u.to_slots(gy, Cp, False, Hp, Wp)
    bs = torch.zeros(2 * Cp, dtype=torch.float64, device='cuda')
    u.call('mmlf_bn_bwd_reduce', u.ptr(gys), Cp, u.ptr(zs), Cp, u.ptr(scale), u.ptr(shift), u.ptr(smean), u.ptr(sinv),
           Cp, B, H, W, G, A, u.ptr(bs), u.stream())
    dz = torch.full((B * Hp * Wp, Cp), float('nan'), dtype=torch.bfloat16, device='cuda')
    dgam, dbet = torch.empty(Cr, device='cuda'), torch.empty(Cr, device='cuda')
    dzsum = torch.zeros(Cp, device='cuda')
    fs = torch.full((3 * Cp,), float('nan'), device='cuda')
    u.call('mmlf_bn_bwd_apply', u.ptr(gys), Cp, u.ptr(zs), Cp, u.ptr(scale), u.ptr(shift), u.ptr(d['gamma']), u.ptr(smean),
           u.ptr(sinv), u.ptr(bs), n, 1, Cr, Cp, B, H, W, G, A, u.ptr(dz), Cp, u.ptr(dgam), u.ptr(dbet), 0,
           u.ptr(fs), u.ptr(dzsum), u.stream())
    # accumulate = 1 (second call of a shared in-net): dgamma / dbeta add to what is there, dz is the same
    dz_b = torch.full_like(dz, float('nan'))
    dzsum2 = torch.zeros(Cp, device='cuda')
    dgam2, dbet2 = dgam.clone(), dbet.clone()
    u.call('mmlf_bn_bwd_apply', u.ptr(gys), Cp, u.ptr(zs), Cp, u.ptr(scale), u.ptr(shift), u.ptr(d['gamma']), u.ptr(smean),
           u.ptr(sinv), u.ptr(bs), n, 1, Cr, Cp, B, H, W, G, A, u.ptr(dz_b), Cp, u.ptr(dgam2), u.ptr(dbet2), 1,
           u.ptr(fs), u.ptr(dzsum2), u.stream())
    torch.cuda.synchronize()
    assert torch.equal(dz, dz_b) and torch.equal(dgam2, 2 * dgam) and torch.equal(dbet2, 2 * dbet)
    assert torch.equal(fs[2 * Cp:2 * Cp + Cr], d['gamma']) and not fs[2 * Cp + Cr:].any()      # gamma on the padded pitch
    np.testing.assert_allclose(dzsum.cpu().numpy(), dz.double().sum(0).cpu().numpy(), rtol=1e-4, atol=1e-3)
    yq = full[:, 1:, 1:, :Cr]
    g = gy * (yq > 0)
    invstd = (1 / np.sqrt(var + 1e-5)).astype(np.float32)
    xhat = ((z - mean) * invstd).astype(np.float32)
    g2, x2 = g.reshape(-1, Cr), xhat.reshape(-1, Cr)
    want_dz = (gamma * invstd * (g - g2.astype(np.float64).mean(0) - xhat * (g2.astype(np.float64) * x2).mean(0))).astype(np.float32)
    g64, x64 = g2.astype(np.float64), x2.astype(np.float64)
    np.testing.assert_allclose(dbet.cpu().numpy(), g64.sum(0), rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(dgam.cpu().numpy(), (g64 * x64).sum(0), rtol=1e-4, atol=2e-3)
    dzf = dz.float().cpu().numpy().reshape(B, Hp, Wp, Cp)
    assert not dzf[:, 0].any() and not dzf[:, :, 0].any()
    u.assert_close_bf16(dzf[:, 1:, 1:, :Cr], want_dz, 'bn_bwd_apply', ulps=1.01, atol=2e-3)


# ------------------------------------------------------------------------------------------------ heads / targets / ESE
def test_heads_against_golden(golden):
    from mmlf_b200 import ops
    g = golden('net_tiny_upr_full.npz')
    mean, logvar = torch.from_numpy(g['eval/mean']).cuda(), torch.from_numpy(g['eval/logvar']).cuda()
    bins = ops.numpy_bins(-3.5, 3.5, 108, 'cuda')
    post = ops.upr_posterior(mean, logvar, bins).cpu().numpy()
    np.testing.assert_allclose(post, g['eval/posterior'], rtol=2e-5, atol=1e-30)
    g = golden('net_tiny_dpp_full.npz')
    s = torch.from_numpy(g['eval/scores']).cuda()
    one_hot, post, m, lv = ops.dpp_head(s, ops.torch_bins(-3.5, 3.5, 108, 'cuda'), bins)
    assert np.array_equal(one_hot.cpu().numpy(), g['eval/one_hot'])
    assert np.array_equal(m.cpu().numpy(), g['eval/mean'])
    np.testing.assert_allclose(post.cpu().numpy(), g['eval/posterior'], rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose(lv.cpu().numpy(), g['eval/logvar'], rtol=1e-4, atol=1e-5)


def test_targets_against_golden(golden):
    from mmlf_b200.utils import dl
    g = golden('bins.npz')
    gt, mpi = torch.from_numpy(g['gt']).cuda(), torch.from_numpy(g['mpi']).cuda()
    for n in (54, 108):
        assert np.array_equal(dl.reg_to_class(gt, -3.5, 3.5, n).cpu().numpy(), g[f'reg_to_class{n}'])
        np.testing.assert_allclose(dl.mpi_to_weights(mpi, -3.5, 3.5, n).cpu().numpy(), g[f'mpi_to_weights{n}'],
                                   rtol=1e-6, atol=1e-7)
        oh = torch.from_numpy(g[f'onehot{n}']).cuda()
        assert np.array_equal(dl.class_to_reg(oh, -3.5, 3.5, n).cpu().numpy(), g[f'class_to_reg{n}'])


def test_ese_reduce_against_golden(golden):
    from mmlf_b200 import ops
    g = golden('ese_tiny.npz')
    for tag in ('coarse', 'full'):
        means, logvars = torch.from_numpy(g[f'{tag}/means']).cuda(), torch.from_numpy(g[f'{tag}/logvars']).cuda()
        disp = ops.numpy_bins(-3.5, 3.5, means.shape[0], 'cuda')
        mean, logvar, post = ops.ese_reduce(means, logvars, disp)
        assert np.array_equal(mean.cpu().numpy(), g[f'{tag}/mean'])
        assert np.array_equal(logvar.cpu().numpy(), g[f'{tag}/logvar'])
        np.testing.assert_allclose(post.cpu().numpy(), g[f'{tag}/posterior'], rtol=3e-5, atol=1e-9)


# ------------------------------------------------------------------------------------------------ losses / Adam
def test_losses_against_golden(golden):
    from mmlf_b200.model import loss as L
    from mmlf_b200.utils import dl
    g = golden('losses.npz')
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    mask, gt, mpi = T(g['mask']), T(g['gt']), T(g['mpi'])

    def run(name, fn, target, *extra, keys=('mean',), m=mask):
        o = {'mean': T(g['mean']).requires_grad_(), 'logvar': T(g['logvar']).requires_grad_(),
             'scores': T(g['scores']).requires_grad_()}
        val = fn(o, target, m, *extra)
        np.testing.assert_allclose(val.item(), float(g[name + '/value']), rtol=1e-5, err_msg=name)
        if keys:
            val.backward()
        for k in keys:
            ref = g[f'{name}/g_{k}']
            np.testing.assert_allclose(o[k].grad.cpu().numpy(), ref, rtol=2e-4, atol=1e-6 * np.abs(ref).max() + 1e-12,
                                       err_msg=name + k)

    run('l1', L.MaskedL1Loss(), gt)
    run('l1_empty', L.MaskedL1Loss(), gt, m=torch.zeros_like(mask))
    run('multi_l1', L.MultiMaskedL1Loss(), mpi)
    run('mse', L.MaskedMSELoss(), gt, keys=())
    run('badpix', L.MaskedBadPix(), gt, keys=())
    run('upr', L.ImprovedUncertaintyL1Loss(), gt, keys=('mean', 'logvar'))
    run('upr_pad', L.ImprovedUncertaintyL1Loss(), gt, T(g['mask_padding']), keys=('mean', 'logvar'))
    run('multi_upr', L.ImprovedMultiUncertaintyL1Loss(), mpi, keys=('mean', 'logvar'))
    run('ce', L.MaskedCrossEntropy(), dl.reg_to_class(gt, -3.5, 3.5, 108), keys=('scores',))
    run('ce_mm', L.MaskedCrossEntropy(), dl.mpi_to_weights(mpi, -3.5, 3.5, 108), keys=('scores',))


def test_ce_on_the_fly_target(golden):
    from mmlf_b200 import ops
    g = golden('losses.npz')
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    mask = T(g['mask'])
    sums = ops.loss_prepass(mask)
    ls, gs = ops.loss_cross_entropy(T(g['scores']), None, T(g['gt']), ops.torch_bins(-3.5, 3.5, 108, 'cuda'),
                                    7.0 / 108 / 2.0, mask, sums)
    np.testing.assert_allclose(ls.item() / sums[0].item(), float(g['ce/value']), rtol=1e-5)
    np.testing.assert_allclose(gs.cpu().numpy(), g['ce/g_scores'], rtol=2e-4, atol=1e-9)


def test_adam_against_golden(golden):
    from mmlf_b200 import ops
    g = golden('adam.npz')
    p = torch.from_numpy(g['p0'].copy()).cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(4):
        ops.adam_step(p, torch.from_numpy(g['grads'][s].copy()).cuda(), m, v, float(g['lrs'][s]), 0.9, 0.999, 1e-8,
                      s + 1)
        np.testing.assert_allclose(p.cpu().numpy(), g[f'p{s + 1}'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m.cpu().numpy(), g['exp_avg'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v.cpu().numpy(), g['exp_avg_sq'], rtol=1e-5, atol=1e-9)


def test_random_shift_against_the_reference(golden):
    """RandomShift (hci4d.py:993-1028) under random.seed(s): the same disparity is drawn from the host `random` stream and
    the resampled stacks / corrected gt / mpi are bit-identical to the reference's numpy branch -- on CUDA tensors and,
    staged through the GPU, on numpy arrays (the reference's dataset-transform usage)."""
    import random
    from mmlf_b200.data import hci4d
    g = golden('randomshift.npz')
    for j, seed in enumerate(g['seeds']):
        lo, hi = (float(x) for x in g['ranges'][j])
        rg = (lo, hi) if bool(g['is_tuple'][j]) else hi
        for numpy_in in (False, True):
            random.seed(int(seed))
            stacks = [g[f'in{k}'].copy() for k in range(4)]
            gt, mpi = g['gt'].copy(), g['mpi'].copy()
            if not numpy_in:
                stacks = [torch.from_numpy(a).cuda() for a in stacks]
                gt, mpi = torch.from_numpy(gt).cuda(), torch.from_numpy(mpi).cuda()
            res = hci4d.RandomShift(rg)(tuple(stacks) + (np.zeros(1), gt, mpi))
            host = lambda t: t if isinstance(t, np.ndarray) else t.cpu().numpy()  # noqa: E731
            for k in range(4):
                assert res[k] is stacks[k]                                          # in place, like the reference
                assert np.array_equal(host(res[k]), g[f'out{j}_{k}']), (j, k, numpy_in)
            assert np.array_equal(host(res[5]), g[f'gt{j}']) and np.array_equal(host(res[6]), g[f'mpi{j}'])
    with pytest.raises(AssertionError):
        hci4d.RandomShift(-1.0)
    with pytest.raises(AssertionError):
        hci4d.RandomShift(1)


def test_hci4d_loader_against_the_oracle(tmp_path):
    """HCI4D.load_scene (hci4d.py:124-254) on an on-disk scene in the dataset's layout: PNG decode on the host, crosshair
    extraction + u8 -> f32 and the texture mask on the GPU; bit-exact against the oracle's extraction of the same bytes."""
    import _fixtures as fx
    import oracle
    from mmlf_b200.data import hci4d
    root = tmp_path / 'training'
    u8a, gta = fx.write_hci_scene(str(root / 'boxes'), 5, 48, 40, with_mask=True, with_mpi=True)
    u8b, gtb = fx.write_hci_scene(str(root / 'cotton'), 6, 48, 40)
    ds = hci4d.HCI4D(str(root), cache=True)
    assert ds.scenes_names == ['boxes', 'cotton'] and len(ds) == 2 and ds.name == 'training'
    assert len(hci4d.HCI4D(str(root), length=4096)) == 4096
    for idx, (u8, gt) in enumerate(((u8a, gta), (u8b, gtb))):
        item = ds[idx]
        want = oracle.extract_stacks(u8)
        for k in range(5):
            assert item[k].is_cuda and np.array_equal(item[k].cpu().numpy(), want[k]), k
        assert np.array_equal(item[5].cpu().numpy(), gt)                            # PFM is stored bottom-up
        assert item[8].tolist() == [idx]
        tex = oracle.create_mask_texture(want[4][None], 23, 0.02)[0]
        sure = np.abs(oracle.texture_mae(want[4][None], 23)[0] - np.float32(0.02)) > 1e-6    # not within round-off of the threshold
        mask = item[7].cpu().numpy()
        assert mask.dtype == np.int64 and mask.shape == gt.shape
        if idx == 0:
            assert not mask[:4].any() and np.array_equal(mask[4:][sure[4:]], tex[4:][sure[4:]])   # mask.png zeroes rows 0..3
            mpi = item[6].cpu().numpy()
            assert mpi.shape == (3, 5, 48, 40) and mpi[2, 4, 3, 5] == 0.0           # NaN -> 0
            assert np.array_equal(mpi[0, 4], gt) and np.array_equal(mpi[1, 3], np.full_like(gt, 0.3))
        else:
            assert np.array_equal(mask[sure], tex[sure])
            mpi = item[6].cpu().numpy()
            assert mpi.shape == (1, 5, 48, 40) and mpi.dtype == np.float64
            assert np.array_equal(mpi[0, :3], want[4]) and np.array_equal(mpi[0, 4], gt) and (mpi[0, 3] == 1).all()
    # the Shift transform works on a copy: the cached scene is untouched (hci4d.py:288-291)
    ds.transform = hci4d.Shift(1.5)
    shifted = ds[1]
    assert np.array_equal(ds.data[1][5].cpu().numpy(), gtb)
    assert np.array_equal(shifted[5].cpu().numpy(), gtb - np.float32(1.5))
    want = oracle.shift(tuple(a.copy() for a in oracle.extract_stacks(u8b)[:4]), 1.5)
    for k in range(4):
        assert np.array_equal(shifted[k].cpu().numpy(), want[k])


@pytest.mark.parametrize('rows', ['', '3', '8'])
def test_pack_rows_per_cta(rows, monkeypatch):
    """The vector packing kernel walks `rows_per_cta` slot rows per CTA (chosen from the problem size; MMLF_PACK_ROWS
    forces it): every choice, including a ragged last group, gives the oracle's bits -- plain and with the fused Shift."""
    u = _u()
    if rows:
        monkeypatch.setenv('MMLF_PACK_ROWS', rows)
    else:
        monkeypatch.delenv('MMLF_PACK_ROWS', raising=False)
    rng = np.random.RandomState(11)
    B, n = 2, 9
    for H, W in ((9, 136), (6, 96), (10, 260)):
        stacks = [rng.uniform(0, 1, (B, n, 3, H, W)).astype(np.float32) for _ in range(4)]
        for disp in (None, 2.5, -3.75):
            sh = stacks if disp is None else oracle.shift(tuple(stacks), disp)
            for k in range(4):
                o = torch.full((B * (H + 1) * (W + 1), 32), float('nan'), dtype=torch.float16, device='cuda')
                sd = torch.from_numpy(stacks[k]).cuda()
                if disp is None:
                    u.call('mmlf_pack_views', u.ptr(sd), B, n * 3, H, W, u.ptr(o), 32, u.FP16, u.stream())
                else:
                    u.call('mmlf_shift_pack', u.ptr(sd), k, B, n, H, W, float(disp), u.ptr(o), 32, u.FP16, u.stream())
                got = o.float().cpu().numpy().reshape(B, H + 1, W + 1, 32)
                want = np.zeros_like(got)
                want[:, 1:, 1:, :27] = u.ROUND[u.FP16](sh[k].reshape(B, 27, H, W).transpose(0, 2, 3, 1))
                assert np.array_equal(got, want), (rows, disp, k, H, W)


@pytest.mark.parametrize('shape', [(2, 9, 20, 24), (1, 9, 17, 19), (3, 5, 32, 32)])
def test_pack_stacks_equals_the_per_stack_launches(shape):
    """mmlf_pack_stacks (all stacks of a forward in one launch, optional second copy in another 16-bit format, optional fused
    ESE Shift) == mmlf_pack_views / mmlf_shift_pack per stack and format, bit for bit."""
    import ctypes as C
    u = _u()
    B, n, H, W = shape
    g = torch.Generator(device='cuda').manual_seed(3)
    views = [torch.rand((B, n, 3, H, W), device='cuda', generator=g) for _ in range(4)]
    ld = u.pad16(n * 3)
    n_slots = B * (H + 1) * (W + 1)
    for disp in (None, 2.5, -0.7):
        for ns in (4, 2):
            outs = [torch.full((n_slots, ld), 7.0, dtype=torch.float16, device='cuda') for _ in range(ns)]
            outs2 = [torch.full((n_slots, ld), 7.0, dtype=torch.bfloat16, device='cuda') for _ in range(ns)]
            PA, IA = C.c_void_p * ns, C.c_int * ns
            u.call('mmlf_pack_stacks', PA(*[v.data_ptr() for v in views[:ns]]), IA(*range(ns)), ns, B, n, H, W,
                   PA(*[t.data_ptr() for t in outs]), PA(*[t.data_ptr() for t in outs2]), ld, u.FP16, u.BF16,
                   0 if disp is None else 1, 0.0 if disp is None else disp, u.stream())
            for si in range(ns):
                for dt, got in ((u.FP16, outs[si]), (u.BF16, outs2[si])):
                    want = torch.full((n_slots, ld), 7.0, dtype=u.TDT[dt], device='cuda')
                    if disp is None:
                        u.call('mmlf_pack_views', u.ptr(views[si]), B, n * 3, H, W, u.ptr(want), ld, dt, u.stream())
                    else:
                        u.call('mmlf_shift_pack', u.ptr(views[si]), si, B, n, H, W, disp, u.ptr(want), ld, dt, u.stream())
                    torch.cuda.synchronize()
                    assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (disp, ns, si, dt)
            # without a second output
            only = [torch.full((n_slots, ld), 7.0, dtype=torch.float16, device='cuda') for _ in range(ns)]
            u.call('mmlf_pack_stacks', PA(*[v.data_ptr() for v in views[:ns]]), IA(*range(ns)), ns, B, n, H, W,
                   PA(*[t.data_ptr() for t in only]), None, ld, u.FP16, u.BF16, 0 if disp is None else 1,
                   0.0 if disp is None else disp, u.stream())
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(only, outs))
