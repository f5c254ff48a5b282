"""The trained-like full-width fixtures (tests/golden/net_trained_*.npz, oracle/gen_golden.py::gen_trained): the reference
model (chs = 70, 4 streams) after real Adam steps -- well-conditioned and non-degenerate, unlike the default-init
checkpoint (near-constant output) and the x2-weights fixtures (chaotic gradients).

CPU part (this file, not gpu): the fp32 oracle reproduces the reference's eval outputs, loss and gradients on them, and
the oracle run with the CUDA path's rounding points ('fp16' emulation) stays inside the bounds the GPU tests assert."""
import numpy as np
import pytest
import torch

import _fixtures as fx
import oracle
from oracle import losses as ol


def build_state(golden, variant):
    """(state, g): parameters rebuilt from the seed + int8 trunk deltas + the variant's stored head / statistics."""
    from mmlf_b200.model.feed_forward import FeedForward
    g = golden(f'net_trained_{variant}.npz')
    trunk = golden('net_trained_trunk.npz')
    torch.manual_seed(0)
    m = FeedForward(**fx.model_kwargs(variant, False, chs=70))
    init = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    return fx.trained_state(variant, init, trunk, g), g


def oracle_loss(variant, e, gt, mask):
    eo = {'mean': e['mean'], 'logvar': e['logvar'], 'scores': e['scores']}
    if variant == 'dpp':
        ev, eg = ol.masked_cross_entropy(eo, ol.reg_to_class(gt, -3.5, 3.5, 108), mask)
        return ev, eg['scores']
    if variant == 'upr':
        ev, eg = ol.improved_uncertainty_l1(eo, gt, mask)
        return ev, np.stack([eg['mean'], eg['logvar']], 1)
    ev, eg = ol.masked_l1(eo, gt, mask)
    return ev, eg['mean'][:, None]


def grad_agreement(grads, g, stride):
    """(cosine over the whole sampled gradient, worst per-tensor relative L2 error, its name).  Conv biases in front of a
    BatchNorm ('.2.bias' of blocks 0..6) have a zero gradient up to round-off and are left out of the per-tensor figure."""
    dot = gg = rr = 0.0
    worst, worst_name = 0.0, ''
    for k in g.files:
        if not k.startswith('grad/'):
            continue
        name = k[5:]
        ref = g[k].astype(np.float64)
        got = np.asarray(grads[name], np.float64)
        if got.shape != ref.shape:
            got = got.reshape(-1)[::stride]
        dot, gg, rr = dot + (got * ref).sum(), gg + (got ** 2).sum(), rr + (ref ** 2).sum()
        bn_fed_bias = name.endswith('.2.bias') and not name.startswith('out_net.7.')
        rel = np.sqrt(((got - ref) ** 2).sum()) / (np.sqrt((ref ** 2).sum()) + 1e-30)
        if not bn_fed_bias and rel > worst:
            worst, worst_name = rel, name
    return dot / np.sqrt(gg * rr), worst, worst_name


@pytest.mark.parametrize('variant', ['base', 'upr', 'dpp'])
def test_oracle_matches_the_reference_on_trained_weights(golden, variant):
    state, g = build_state(golden, variant)
    (h, v, i, d), gt, mask = fx.trained_batch(0)
    o = oracle.FeedForwardOracle(state, model_uncert=variant == 'upr', model_discrete=variant == 'dpp')
    key = 'scores' if variant == 'dpp' else 'mean'
    e = o.forward(h, v, i, d)
    got = e[key][[0, 5]] if variant == 'dpp' else e[key]
    ref = g['eval/' + key]
    rng = float(ref.max() - ref.min())
    assert rng > 1.0, 'fixture output is degenerate'
    assert np.abs(got - ref).max() <= 2e-6 * rng
    if variant == 'upr':
        assert np.abs(e['logvar'] - g['eval/logvar']).max() <= 5e-6
    o.training = True
    e = o.forward(h, v, i, d, keep_tape=True)
    ev, gout = oracle_loss(variant, e, gt, mask)
    assert abs(ev - float(g['train/loss'])) <= 1e-6 * abs(float(g['train/loss']))
    cos, worst, name = grad_agreement(o.backward(gout), g, fx.TRAINED['grad_stride'])
    assert cos >= 0.999999 and worst <= 2e-3, (cos, worst, name)


def test_emulated_cuda_rounding_stays_inside_the_gpu_bounds(golden):
    """The oracle with the kernels' rounding points (fp16 activations / weights, bf16 gradients, fp32 accumulate) on the
    UPR fixture: this is what the GPU parity test (tests/test_gpu_model.py::test_trained_like_models) can achieve."""
    state, g = build_state(golden, 'upr')
    (h, v, i, d), gt, mask = fx.trained_batch(0)
    o = oracle.FeedForwardOracle(state, model_uncert=True, quant='fp16')
    e = o.forward(h, v, i, d)
    ref = g['eval/mean']
    assert np.abs(e['mean'] - ref).max() <= 5e-3 * float(ref.max() - ref.min())
    o.training = True
    e = o.forward(h, v, i, d, keep_tape=True)
    ev, gout = oracle_loss('upr', e, gt, mask)
    assert abs(ev - float(g['train/loss'])) <= 1e-3 * abs(float(g['train/loss']))
    cos, worst, name = grad_agreement(o.backward(gout), g, fx.TRAINED['grad_stride'])
    assert cos >= 0.999 and worst <= 0.05, (cos, worst, name)


def test_trained_trunk_is_not_the_initialisation(golden):
    state, g = build_state(golden, 'base')
    from mmlf_b200.model.feed_forward import FeedForward
    torch.manual_seed(0)
    init = FeedForward(**fx.model_kwargs('base', False, chs=70)).state_dict()
    moved = [float(np.abs(state[k] - init[k].numpy()).max()) for k in state if k.endswith('.weight') and state[k].ndim == 4]
    assert min(moved) > 1e-3 and max(moved) < 0.1             # 30 Adam steps at lr 1e-3
    traj = g['traj/loss']
    assert traj.shape == (fx.TRAINED['traj_steps'],) and np.isfinite(traj).all()
