"""Parity at BASELINE.json's full sizes through size-independent properties (-m gpu), and the degenerate shapes.

The oracle finishes in seconds only on small inputs, so at the benchmark sizes (64 patches of 96 px = the per-GPU share of
the 512-patch batch at 8 GPUs; one 9x9-view 512x512 light field) the kernels are checked through properties that do not
need it: an integer disparity shift is a circular roll, the tensor-core convolution equals the CUDA-core cross-check
kernel, the statistics fused into the conv epilogue equal the stand-alone reduction kernel, a loss over the whole batch
equals the sum over its shards, the ESE posterior integrates to one, Adam leaves zero-gradient parameters alone.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
import _gpu_util as u

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ light field
@pytest.mark.parametrize('shape', [(1, 9, 512, 512), (64, 9, 96, 96)])
def test_integer_shift_is_a_roll_at_full_size(shape):
    """Shift(1.0): alpha = 0 for every view, so view k of the h stack is rolled by (k - 4) pixels along W, of the v stack
    along H, i: +W then -H, d: +W then +H (hci4d.py:934-981) -- bit exact, any size."""
    from mmlf_b200 import ops
    B, n, H, W = shape
    g = torch.Generator(device='cuda').manual_seed(0)
    stacks = [torch.rand((B, n, 3, H, W), device='cuda', generator=g) for _ in range(4)]
    out = ops.lf_shift(*stacks, 1.0)
    for k in range(n):
        s = k - n // 2
        assert torch.equal(out[0][:, k], torch.roll(stacks[0][:, k], s, -1))
        assert torch.equal(out[1][:, k], torch.roll(stacks[1][:, k], s, -2))
        assert torch.equal(out[2][:, k], torch.roll(torch.roll(stacks[2][:, k], s, -1), -s, -2))
        assert torch.equal(out[3][:, k], torch.roll(torch.roll(stacks[3][:, k], s, -1), s, -2))
    # disparity 0 is the identity
    same = ops.lf_shift(*stacks, 0.0)
    assert all(torch.equal(a, b) for a, b in zip(same, stacks))


def test_shift_is_linear_in_the_input_at_full_size():
    """The resampler is a fixed linear map: shift(a) + shift(b) == shift(a + b) up to the fp32 rounding of the lerp."""
    from mmlf_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(1)
    a = [torch.rand((1, 9, 3, 512, 512), device='cuda', generator=g) for _ in range(4)]
    b = [torch.rand((1, 9, 3, 512, 512), device='cuda', generator=g) for _ in range(4)]
    sa, sb = ops.lf_shift(*a, 2.5), ops.lf_shift(*b, 2.5)
    sab = ops.lf_shift(*[x + y for x, y in zip(a, b)], 2.5)
    for x, y, z in zip(sa, sb, sab):
        assert (x + y - z).abs().max().item() < 1e-6


def test_pack_views_roundtrip_at_full_size():
    """fp16 slots of a full light field stack: every pixel lands at slot (y + 1, x + 1), halo and padding channels zero."""
    g = torch.Generator(device='cuda').manual_seed(2)
    v = torch.rand((1, 9, 3, 512, 512), device='cuda', generator=g)
    out = torch.full((513 * 513, 32), float('nan'), dtype=torch.float16, device='cuda')
    u.call('mmlf_pack_views', u.ptr(v), 1, 27, 512, 512, u.ptr(out), 32, u.FP16, u.stream())
    s = out.view(513, 513, 32)
    assert torch.equal(s[1:, 1:, :27], v.view(27, 512, 512).permute(1, 2, 0).to(torch.float16))
    assert not s[0].any() and not s[:, 0].any() and not s[..., 27:].any()


# ------------------------------------------------------------------------------------------------ convolution
@pytest.mark.parametrize('case', [(64, 96, 96, 280, 280, 1), (1, 512, 512, 280, 280, 0), (64, 96, 96, 70, 70, 0)])
def test_tensor_core_conv_equals_cuda_core_conv_at_full_size(case):
    """Same operands through tcgen05 (TMA, TMEM, fused epilogue) and through the fp32-FMA cross-check kernel."""
    B, H, W, cin, cout, ctype = case
    cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
    n_slots = B * (H + 1) * (W + 1)
    g = torch.Generator(device='cuda').manual_seed(3)
    x = (torch.randn((n_slots, cin_pad), device='cuda', generator=g) * 0.5).to(torch.float16)
    x[:, cin:] = 0
    w = (np.random.RandomState(0).normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    wp = u.pack_weight(w, dt=u.FP16)
    bias = torch.zeros(n_pad, device='cuda')
    bias[:cout] = torch.linspace(-0.2, 0.2, cout)
    tc = u.run_conv(x, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, bias=bias, relu=True, ab=u.FP16, out_dt=u.FP16)
    ref = u.run_conv(x, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, bias=bias, relu=True, ab=u.FP16, out_dt=u.FP16,
                     simt=True)
    d = (tc.float() - ref.float()).abs()
    # both accumulate in fp32 (different order) and round once to fp16: at most one fp16 ulp apart
    tol = ref.float().abs() * 2.0 ** -10 + 1e-4
    assert bool((d <= tol).all()), float((d - tol).max())
    assert float((d > 0).float().mean()) < 0.05                 # and bit-identical almost everywhere


def test_epilogue_statistics_equal_the_reduction_kernel_at_full_size():
    """BatchNorm batch statistics: conv epilogue (per-warp fp32 partials + fp64 atomics) vs the stand-alone column
    reduction over the stored activations, 64 patches x 96 px x 280 channels."""
    B, H, W, cin, cout = 64, 96, 96, 280, 280
    cin_pad = n_pad = u.pad16(cin)
    n_slots = B * (H + 1) * (W + 1)
    g = torch.Generator(device='cuda').manual_seed(4)
    x = (torch.randn((n_slots, cin_pad), device='cuda', generator=g) * 0.5).to(torch.float16)
    w = (np.random.RandomState(1).normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
    wp = u.pack_weight(w, dt=u.FP16)
    fused = torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda')
    z = u.run_conv(x, cin_pad, cin_pad, wp, n_pad, B, H, W, 1, ab=u.FP16, out_dt=u.FP16, col_sums=fused)
    alone = torch.zeros(2 * n_pad, dtype=torch.float64, device='cuda')
    u.call('mmlf_bn_stats', u.ptr(z), n_pad, n_pad, B, H, W, u.FP16, u.ptr(alone), u.stream())
    torch.cuda.synchronize()
    ref = z.double()
    want = torch.cat([ref.sum(0), (ref * ref).sum(0)])
    assert torch.allclose(alone, want, rtol=1e-6, atol=1e-3)
    assert torch.allclose(fused, want, rtol=2e-5, atol=5e-2)


# ------------------------------------------------------------------------------------------------ losses / heads
def test_loss_is_additive_over_batch_shards_at_full_size():
    """sum-type kernels: loss over 64 patches == sum over 8 shards of 8 with the same global normalisers (this is the
    multi-GPU protocol of SURVEY.md H4)."""
    from mmlf_b200 import ops
    B, H, W = 64, 96, 96
    g = torch.Generator(device='cuda').manual_seed(5)
    mean = torch.randn((B, H, W), device='cuda', generator=g)
    logvar = torch.randn((B, H, W), device='cuda', generator=g) * 0.3
    gt = torch.randn((B, H, W), device='cuda', generator=g)
    mask = (torch.rand((B, H, W), device='cuda', generator=g) > 0.2).to(torch.int32)
    sums = ops.loss_prepass(mask)
    assert sums[0].item() == mask.sum().item() and sums[4].item() == B * H * W
    whole, gm, gl = ops.loss_regression(2, mean, logvar, gt, mask, None, sums)
    parts = 0.0
    for s in range(0, B, 8):
        sl = slice(s, s + 8)
        p, pgm, pgl = ops.loss_regression(2, mean[sl].contiguous(), logvar[sl].contiguous(), gt[sl].contiguous(),
                                          mask[sl].contiguous(), None, sums)
        parts += p.item()
        assert torch.equal(pgm, gm[sl]) and torch.equal(pgl, gl[sl])
    assert abs(parts - whole.item()) < 1e-9 * abs(whole.item()) + 1e-9
    ref = ((torch.exp(-logvar.double()) * (mean.double() - gt.double()).abs() + logvar.double()) * mask).sum().item()
    assert abs(whole.item() - ref) < 1e-5 * abs(ref)


def test_shared_memory_cross_entropy_equals_the_generic_kernel():
    """loss_ce: the shared-memory kernel (HW % 128 == 0, 16-byte aligned inputs) gives bit-identical gradients and the same
    value as the generic kernel, for the on-the-fly and the dense target; a copy of the scores at a 4-byte offset takes
    the generic path."""
    from mmlf_b200 import ops
    from mmlf_b200.utils import dl
    g = torch.Generator(device='cuda').manual_seed(8)
    B, S, H, W = 8, 108, 96, 96
    scores = torch.randn((B, S, H, W), device='cuda', generator=g) * 2
    gt = torch.rand((B, H, W), device='cuda', generator=g) * 6 - 3
    mask = (torch.rand((B, H, W), device='cuda', generator=g) > 0.2).int()
    bins_t = ops.torch_bins(-3.5, 3.5, S, 'cuda')
    sums = ops.loss_prepass(mask)
    buf = torch.empty(scores.numel() + 1, device='cuda')
    shifted = buf[1:].view(scores.shape)
    shifted.copy_(scores)
    dense = dl.reg_to_class(gt, -3.5, 3.5, S)
    for tgt, gt_, bins in ((None, gt, bins_t), (dense, None, None)):
        l1, g1 = ops.loss_cross_entropy(scores, tgt, gt_, bins, 7.0 / S / 2.0, mask, sums)
        l2, g2 = ops.loss_cross_entropy(shifted, tgt, gt_, bins, 7.0 / S / 2.0, mask, sums)
        assert torch.equal(g1, g2)
        assert abs(l1.item() - l2.item()) <= 1e-12 * abs(l2.item())
    # on-the-fly target == dense target
    assert torch.equal(g1, ops.loss_cross_entropy(scores, None, gt, bins_t, 7.0 / S / 2.0, mask, sums)[1])


def test_posteriors_are_normalised_at_full_size():
    """DPP softmax posterior sums to 1 per pixel (108 bins, 64 x 96 x 96); the one-hot marks the arg-max; the ESE Laplace
    mixture is non-negative and its mean/logvar are members of the ensemble."""
    from mmlf_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(6)
    scores = torch.randn((64, 108, 96, 96), device='cuda', generator=g) * 2
    bins_t, bins_n = ops.torch_bins(-3.5, 3.5, 108, 'cuda'), ops.numpy_bins(-3.5, 3.5, 108, 'cuda')
    one_hot, post, mean, logvar = ops.dpp_head(scores, bins_t, bins_n)
    assert (post.sum(1) - 1).abs().max().item() < 1e-5
    assert torch.equal(one_hot.argmax(1), scores.argmax(1)) and one_hot.sum().item() == 64 * 96 * 96
    assert torch.equal(mean, bins_t[scores.argmax(1)])
    # the shared-memory kernel (HW % 128 == 0, 16-byte aligned scores) == the generic kernel, bit for bit: a copy of the
    # scores at a 4-byte offset takes the generic path
    buf = torch.empty(scores.numel() + 1, device='cuda')
    shifted = buf[1:].view(scores.shape)
    shifted.copy_(scores)
    assert shifted.data_ptr() % 16 == 4
    for a, b in zip((one_hot, post, mean, logvar), ops.dpp_head(shifted, bins_t, bins_n)):
        assert torch.equal(a, b)
    K = 70
    means = torch.randn((K, 1, 512, 512), device='cuda', generator=g)
    logvars = torch.randn((K, 1, 512, 512), device='cuda', generator=g) * 0.3
    m, lv, p = ops.ese_reduce(means, logvars, ops.numpy_bins(-3.5, 3.5, K, 'cuda'))
    idx = logvars.argmin(0)
    assert torch.equal(lv, logvars.gather(0, idx[None])[0]) and torch.equal(m, means.gather(0, idx[None])[0])
    assert p.min().item() >= 0 and torch.isfinite(p).all()


def test_adam_leaves_zero_gradient_parameters_alone_at_full_size():
    from mmlf_b200 import ops
    n = 4612166
    g = torch.Generator(device='cuda').manual_seed(7)
    p = torch.randn(n, device='cuda', generator=g)
    p0 = p.clone()
    grad = torch.zeros(n, device='cuda')
    grad[::2] = torch.randn(n // 2, device='cuda', generator=g)
    m, v = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    ops.adam_step(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 1)
    assert torch.equal(p[1::2], p0[1::2])
    # first step of Adam moves every other parameter by lr * sign(g) (up to eps)
    moved = (p[::2] - p0[::2])
    big = grad[::2].abs() > 1e-3                              # |g| >> eps; p - p0 itself carries an fp32 rounding of ~2e-7
    assert torch.allclose(moved[big], -1e-3 * torch.sign(grad[::2][big]), rtol=2e-3, atol=5e-7)


# ------------------------------------------------------------------------------------------------ degenerate shapes
@pytest.mark.parametrize('shape', [(1, 1, 1), (1, 1, 7), (3, 2, 1), (1, 5, 3)])
def test_conv_on_degenerate_images(shape):
    """1-pixel images, single rows / columns: every tile is partial and every tap touches the zero halo."""
    from oracle.net import conv2x2
    B, H, W = shape
    rng = np.random.RandomState(8)
    for cin, cout, ctype in ((27, 70, 0), (70, 70, 1), (280, 280, 0)):
        Hp, Wp = H + 1, W + 1
        cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
        w = (rng.normal(0, 1, (cout, cin, 2, 2)) / np.sqrt(4 * cin)).astype(np.float32)
        b = rng.normal(0, 0.1, cout).astype(np.float32)
        x = rng.normal(0, 1, (B, H, W, cin) if ctype == 0 else (B, Hp, Wp, cin)).astype(np.float32)
        xq, wq = u.ROUND[u.FP16](x), u.ROUND[u.FP16](w)
        want = np.maximum(conv2x2(xq, wq, b, 1 if ctype == 0 else 0), 0)
        xs = u.to_slots(xq, cin_pad, ctype == 1, Hp, Wp, u.FP16)
        wp = u.pack_weight(w, dt=u.FP16)
        bias = torch.zeros(n_pad, device='cuda')
        bias[:cout] = torch.from_numpy(b)
        out = u.run_conv(xs, cin_pad, cin_pad, wp, n_pad, B, H, W, ctype, bias=bias, relu=True, ab=u.FP16, out_dt=u.FP16)
        full = out.float().cpu().numpy().reshape(B, Hp, Wp, -1)
        got = full[..., :cout] if ctype == 0 else full[:, 1:, 1:, :cout]
        u.assert_close_bf16(got, want, f'conv {shape} {cin}->{cout}', ulps=1.01, atol=3e-4, dt=u.FP16)


def test_model_on_a_tiny_image_and_single_member_ensemble(golden):
    """A 3 x 3 pixel light field through the whole network, and an ensemble of one member."""
    import _fixtures as fx
    from mmlf_b200.model.ensamble import Ensamble
    from mmlf_b200.model.feed_forward import FeedForward
    g = golden('net_tiny_upr_full.npz')
    kw = fx.model_kwargs('upr', False, chs=8)
    m = FeedForward(**kw)
    m.load_state_dict({k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith('state/')}, strict=False)
    m = m.cuda().eval()
    state = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    h, v, i, d, gt = fx.synth_batch(5, 1, 3, 3)
    T = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    with torch.no_grad():
        out = m(T(h), T(v), T(i), T(d))
        emu = oracle.FeedForwardOracle(state, model_uncert=True, quant='fp16').forward(h, v, i, d)
        scale = max(float(np.abs(emu['mean']).max()), 1e-3)
        assert np.abs(out['mean'].cpu().numpy() - emu['mean']).max() <= 0.03 * scale + 1e-3
        ens = Ensamble(m, -3.5, 3.5, 7.0)                      # np.arange(-3.5, 3.5, 7.0) = [-3.5]: one member
        r = ens(T(h), T(v), T(i), T(d))
        assert r['means'].shape[0] == 1 and torch.equal(r['mean'], r['means'][0]) and torch.equal(r['logvar'], r['logvars'][0])


# ------------------------------------------------------------------------------------------------ oracle at BASELINE size
@pytest.mark.parametrize('case', [(1, 96, 96, 280, 280, 0), (1, 96, 96, 280, 280, 1), (1, 512, 512, 280, 280, 1),
                                  (1, 512, 512, 280, 280, 0)])
def test_tensor_core_conv_against_the_oracle_at_baseline_size(case):
    """The numpy oracle can do ONE image at the benchmark sizes: a 96 x 96 patch and a 512 x 512 light-field tile of the
    280 -> 280 out-net convolutions (both block positions), tcgen05 kernel vs oracle within one ulp of the fp16 storage
    (the BatchNorm passes at this size: tests/test_gpu_kernels.py::test_bn_train_roundtrip[shape1])."""
    from test_gpu_kernels import _conv_case
    _conv_case(*case, seed=21, simt=False, dt=1)
