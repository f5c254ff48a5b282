"""SURVEY.md section 8f.4 on the GPU (-m gpu): odd --model_ksize and the --model_unet out-net, run by
mmlf_b200.engine_generic.GenericEngine on the float32 layer kernels of csrc/generic.cu.  Kernel-level checks against the
oracle's layer functions, model-level checks against the REFERENCE's outputs (tests/golden/net_tiny_base_k3.npz,
net_unet_upr.npz: eval outputs, training loss, every parameter gradient, BatchNorm statistics)."""
import ctypes as C

import numpy as np
import pytest
import torch

import _fixtures as fx
from oracle import net as onet
from oracle import unet as ounet

pytestmark = pytest.mark.gpu
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731


def _call(name, *args):
    from mmlf_b200._lib import call
    call(name, *args)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _transform(x, spatial, inverse=False):
    """The data-side stream plumbing of feed_forward.py:236-256 on NHWC arrays: 1 = swap H / W, 2 = swap then flip W."""
    if spatial == 0:
        return x
    if spatial == 1:
        return x.transpose(0, 2, 1, 3)
    if not inverse:
        return x.transpose(0, 2, 1, 3)[:, :, ::-1, :]
    return x[:, :, ::-1, :].transpose(0, 2, 1, 3)


@pytest.mark.parametrize('case', [(2, 9, 11, 5, 7, 3, 1, 0), (1, 8, 8, 27, 70, 3, 1, 1), (2, 7, 7, 6, 9, 3, 1, 2),
                                  (2, 6, 9, 12, 10, 2, 1, 0), (2, 6, 6, 12, 10, 2, 0, 2), (1, 10, 9, 4, 130, 5, 2, 1),
                                  (3, 5, 6, 70, 3, 1, 0, 0), (1, 33, 31, 64, 64, 3, 1, 0)])
def test_generic_conv_forward_backward(case):
    """mmlf_g_conv / _g_pack_weight / _g_conv_wgrad / _g_colsum against the oracle's conv (any k, pad) -- the `spatial`
    tap mapping against the explicit transposes / flips of the data."""
    B, H, W, cin, cout, k, pad, spatial = case
    rng = np.random.RandomState(5)
    x = rng.normal(0, 1, (B, H, W, cin)).astype(np.float32)
    w = rng.normal(0, 0.3, (cout, cin, k, k)).astype(np.float32)
    b = rng.normal(0, 0.3, cout).astype(np.float32)
    xt = np.ascontiguousarray(_transform(x, spatial))
    want_t = onet.conv2x2(xt, w, b, pad)
    want = np.ascontiguousarray(_transform(want_t, spatial, inverse=True))
    Ho, Wo = want.shape[1:3]
    gy = rng.normal(0, 1, want.shape).astype(np.float32)
    gx_t, gw, gb = onet.conv2x2_bwd(xt, w, np.ascontiguousarray(_transform(gy, spatial)), pad)
    gx = np.ascontiguousarray(_transform(gx_t, spatial, inverse=True))
    xd, wd, bd, gyd = T(x.reshape(-1, cin)), T(w), T(b), T(gy.reshape(-1, cout))
    wf = torch.empty(w.size, device='cuda')
    _call('mmlf_g_pack_weight', _p(wd), cout, cin, k, spatial, 0, _p(wf), _st())
    y = torch.full((B * Ho * Wo, cout + 3), float('nan'), device='cuda')              # pitch > C
    _call('mmlf_g_conv', _p(xd), cin, _p(wf), _p(bd), B, H, W, cin, cout, k, pad, 0, _p(y), cout + 3, _st())
    got = y[:, :cout].cpu().numpy().reshape(want.shape)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max())
    assert torch.isnan(y[:, cout:]).all()
    _call('mmlf_g_conv', _p(xd), cin, _p(wf), _p(bd), B, H, W, cin, cout, k, pad, 1, _p(y), cout + 3, _st())
    np.testing.assert_allclose(y[:, :cout].cpu().numpy().reshape(want.shape), np.maximum(want, 0), rtol=2e-5,
                               atol=2e-5 * np.abs(want).max())
    wb = torch.empty(w.size, device='cuda')
    _call('mmlf_g_pack_weight', _p(wd), cout, cin, k, spatial, 1, _p(wb), _st())
    gxd = torch.full((B * H * W, cin), float('nan'), device='cuda')
    _call('mmlf_g_conv', _p(gyd), cout, _p(wb), _p(None), B, Ho, Wo, cout, cin, k, k - 1 - pad, 0, _p(gxd), cin, _st())
    np.testing.assert_allclose(gxd.cpu().numpy().reshape(gx.shape), gx, rtol=2e-5, atol=2e-5 * np.abs(gx).max())
    dw = torch.zeros_like(wd)
    db = torch.zeros(cout, device='cuda')
    for rep in range(2):                                                                 # adds: twice = 2 x
        _call('mmlf_g_conv_wgrad', _p(xd), cin, _p(gyd), cout, B, H, W, cin, cout, k, pad, spatial, 0, _p(dw), _st())
        _call('mmlf_g_colsum', _p(gyd), cout, cout, B * Ho * Wo, _p(db), _st())
    np.testing.assert_allclose(dw.cpu().numpy(), 2 * gw, rtol=1e-4, atol=1e-4 * np.abs(gw).max())
    np.testing.assert_allclose(db.cpu().numpy(), 2 * gb, rtol=1e-4, atol=1e-4 * np.abs(gb).max())


def test_generic_pool_upconv_batchnorm():
    rng = np.random.RandomState(9)
    B, H, W, Cc = 2, 7, 10, 6                                                            # odd H: the last row is dropped
    x = rng.normal(0, 1, (B, H, W, Cc)).astype(np.float32)
    x[0, 0, 0, 0] = x[0, 0, 1, 0] = 5.0                                                  # a tie: the first maximum wins
    want, rec = ounet._pool_fwd(x)
    xd = T(x.reshape(-1, Cc))
    y = torch.empty((B * (H // 2) * (W // 2), Cc), device='cuda')
    idx = torch.empty(y.numel(), dtype=torch.uint8, device='cuda')
    _call('mmlf_g_maxpool2', _p(xd), B, H, W, Cc, _p(y), _p(idx), _st())
    assert np.array_equal(y.cpu().numpy().reshape(want.shape), want)
    gy = rng.normal(0, 1, want.shape).astype(np.float32)
    gx = torch.full((B * H * W, Cc), float('nan'), device='cuda')
    gyd = T(gy.reshape(-1, Cc))
    _call('mmlf_g_maxpool2_bwd', _p(gyd), _p(idx), B, H, W, Cc, _p(gx), _st())
    assert np.array_equal(gx.cpu().numpy().reshape(x.shape), ounet._pool_bwd(rec, gy))
    # transposed conv = 1x1 conv to 4 * cout + depth-to-space
    cin, cu, h, w_ = 5, 3, 4, 6
    xu = rng.normal(0, 1, (B, h, w_, cin)).astype(np.float32)
    wu = rng.normal(0, 0.5, (cin, cu, 2, 2)).astype(np.float32)
    bu = rng.normal(0, 0.5, cu).astype(np.float32)
    want = ounet._upconv_fwd(xu, wu, bu)
    wf = torch.empty(wu.size, device='cuda')
    wud, b4d = T(wu), T(np.tile(bu, 4))
    _call('mmlf_g_pack_weight', _p(wud), cu, cin, 2, 0, 2, _p(wf), _st())
    y4 = torch.empty((B * h * w_, 4 * cu), device='cuda')
    xud = T(xu.reshape(-1, cin))
    _call('mmlf_g_conv', _p(xud), cin, _p(wf), _p(b4d), B, h, w_, cin, 4 * cu, 1, 0, 0, _p(y4), 4 * cu, _st())
    cat = torch.full((B * 2 * h * 2 * w_, cu + 2), float('nan'), device='cuda')
    _call('mmlf_g_depth_to_space', _p(y4), _p(cat), cu + 2, 0, B, h, w_, cu, 0, _st())
    np.testing.assert_allclose(cat[:, :cu].cpu().numpy().reshape(want.shape), want, rtol=1e-5, atol=1e-5)
    gout = rng.normal(0, 1, want.shape).astype(np.float32)
    gx_w, gw_w, gb_w = ounet._upconv_bwd(xu, wu, gout)
    gcat = T(gout.reshape(-1, cu))
    g4 = torch.empty_like(y4)
    _call('mmlf_g_depth_to_space', _p(g4), _p(gcat), cu, 0, B, h, w_, cu, 1, _st())
    dw = torch.zeros((cin, cu, 2, 2), device='cuda')
    _call('mmlf_g_conv_wgrad', _p(xud), cin, _p(g4), 4 * cu, B, h, w_, cin, 4 * cu, 1, 0, 0, 1, _p(dw), _st())
    np.testing.assert_allclose(dw.cpu().numpy(), gw_w, rtol=1e-4, atol=1e-5)
    wb = torch.empty(wu.size, device='cuda')
    _call('mmlf_g_pack_weight', _p(wud), cu, cin, 2, 0, 3, _p(wb), _st())
    gxd = torch.empty((B * h * w_, cin), device='cuda')
    _call('mmlf_g_conv', _p(g4), 4 * cu, _p(wb), _p(None), B, h, w_, 4 * cu, cin, 1, 0, 0, _p(gxd), cin, _st())
    np.testing.assert_allclose(gxd.cpu().numpy().reshape(gx_w.shape), gx_w, rtol=1e-4, atol=1e-5)
    # BatchNorm backward (U-Net order: no gate) against the oracle
    a = rng.normal(0.2, 1.3, (B, H, W, Cc)).astype(np.float32)
    p = {'bn.weight': rng.uniform(0.5, 1.5, Cc).astype(np.float32), 'bn.bias': np.zeros(Cc, np.float32),
         'bn.running_mean': np.zeros(Cc, np.float32), 'bn.running_var': np.ones(Cc, np.float32),
         'bn.num_batches_tracked': np.zeros((), np.int64)}
    yb, brec = ounet._bn_fwd(p, 'bn', a, True)
    gyb = rng.normal(0, 1, a.shape).astype(np.float32)
    gx_want, gr = ounet._bn_bwd(p, brec, gyb, True)
    mean = a.reshape(-1, Cc).astype(np.float64).mean(0).astype(np.float32)
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device='cuda')
    dgam, dbet = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
    gxb = torch.empty((B * H * W, Cc), device='cuda')
    keep = [T(gyb.reshape(-1, Cc)), T(a.reshape(-1, Cc)), T(p['bn.weight']), T(mean), T(brec['invstd'])]   # stay alive
    _call('mmlf_g_bn_bwd', _p(keep[0]), Cc, _p(keep[1]), Cc, _p(None), 0, _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(sums),
          B * H * W, 1, Cc, B * H * W, _p(gxb), Cc, _p(dgam), _p(dbet), _st())
    np.testing.assert_allclose(gxb.cpu().numpy().reshape(a.shape), gx_want, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dgam.cpu().numpy(), gr['bn.weight'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dbet.cpu().numpy(), gr['bn.bias'], rtol=1e-4, atol=1e-4)


def _grad_agreement(m, g, stride=97):
    dot = gg = rr = 0.0
    worst, worst_name = 0.0, ''
    for name, p in m.named_parameters():
        ref = g['grad/' + name].astype(np.float64)
        got = p.grad.cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all(), name
        if got.shape != ref.shape:
            got = got.reshape(-1)[::stride]
        dot, gg, rr = dot + (got * ref).sum(), gg + (got ** 2).sum(), rr + (ref ** 2).sum()
        rel = np.sqrt(((got - ref) ** 2).sum()) / (np.sqrt((ref ** 2).sum()) + 1e-30)
        zero_grad_bias = ('.2.bias' in name and not name.startswith('out_net.7.')) or name.endswith('block.0.bias') or \
            name.endswith('block.3.bias')                       # conv biases in front of a BatchNorm: noise only
        if not zero_grad_bias and rel > worst:
            worst, worst_name = rel, name
    return dot / np.sqrt(gg * rr), worst, worst_name


def test_ksize3_model_against_the_reference(golden):
    from mmlf_b200.model import loss as L
    from mmlf_b200.model.feed_forward import FeedForward
    g = golden('net_tiny_base_k3.npz')
    m = FeedForward(**fx.model_kwargs('base', False, chs=8, model_ksize=3))
    m.load_state_dict({k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith('state/')})
    m = m.cuda()
    h, v, i, d, gt = fx.synth_batch(21, 2, 20, 20)
    mask = fx.synth_mask(22, 2, 20, 20)
    args = [T(a) for a in (h, v, i, d)]
    m.eval()
    with torch.no_grad():
        out = m(*args)
        again = m(*args)                                           # replayed from the captured inference graph
    ref = g['eval/mean']
    err = float(np.abs(out['mean'].cpu().numpy() - ref).max() / (ref.max() - ref.min()))
    assert err <= 2e-4, err
    assert torch.equal(out['mean'], again['mean'])
    m.train()
    out = m(*args)
    lossv = L.MaskedL1Loss()(out, T(gt), T(mask))
    lossv.backward()
    assert abs(lossv.item() - float(g['train/loss'])) <= 1e-4 * abs(float(g['train/loss']))
    np.testing.assert_allclose(out['mean'].detach().cpu().numpy(), g['train/mean'], rtol=0, atol=5e-4 * np.abs(g['train/mean']).max())
    cos, worst, name = _grad_agreement(m, g)
    assert cos >= 0.9995 and worst <= 0.05, (cos, worst, name)
    for k in g.files:
        if k.startswith('after/'):
            np.testing.assert_allclose(m.state_dict()[k[6:]].cpu().numpy(), g[k], rtol=1e-4, atol=1e-6, err_msg=k)


def test_unet_model_against_the_reference(golden):
    """--model_unet: UPR model with the U-Net out-net (31 M parameters, regenerated from the seed like the generator did)."""
    from mmlf_b200.model import loss as L
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    g = golden('net_unet_upr.npz')
    shapes = [(str(n), tuple(int(x) for x in str(s).split(',')) if str(s) else ()) for n, s in zip(g['names'], g['shapes'])]
    state = fx.synth_state(shapes, 17)
    m = FeedForward(**fx.model_kwargs('upr', False, chs=8, model_unet=True))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    m = m.cuda()
    h, v, i, d, gt = fx.synth_batch(51, 2, 32, 32)
    mask = fx.synth_mask(52, 2, 32, 32)
    args = [T(a) for a in (h, v, i, d)]
    m.eval()
    with torch.no_grad():
        out = m(*args)
    for k in ('mean', 'logvar'):
        ref = g['eval/' + k]
        assert np.abs(out[k].cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), k
    assert out['posterior'].shape == (2, m.steps, 32, 32)
    m.train()
    out = m(*args)
    lossv = L.ImprovedUncertaintyL1Loss()(out, T(gt), T(mask))
    lossv.backward()
    for k in ('mean', 'logvar'):
        ref = g['train/' + k]
        assert np.abs(out[k].detach().cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max(), k
    np.testing.assert_allclose(lossv.item(), float(g['train/loss']), rtol=1e-3)
    cos, worst, name = _grad_agreement(m, g)
    assert cos >= 0.999, (cos, worst, name)
    for k in g.files:
        if k.startswith('after/'):
            np.testing.assert_allclose(m.state_dict()[k[6:]].cpu().numpy(), g[k], rtol=1e-3, atol=1e-5, err_msg=k)
    # the captured training step drives this engine too: two steps, the second a graph replay, loss goes down
    opt = FusedAdam(m.parameters(), lr=1e-4)
    step = TrainStep(m, opt, 'upr')
    l0 = step(*args, T(gt), T(mask)).item()
    l1 = step(*args, T(gt), T(mask)).item()
    l2 = step(*args, T(gt), T(mask)).item()
    assert step.replays == 2 and np.isfinite([l0, l1, l2]).all() and l2 < l0
