"""The product's training step (-m gpu): mmlf_b200.train.step.TrainStep -- forward + loss + backward + Adam as one
CUDA-graph replay, no autograd, no framework kernels -- against the eager autograd path and against the reference's own
Adam trajectory; graph / weight-pack invalidation when a parameter is written behind the step's back."""
import copy

import numpy as np
import pytest
import torch

import _fixtures as fx

pytestmark = pytest.mark.gpu
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731


def _model(variant, golden=None, chs=8, seed=3):
    """With `golden`: the trained-like full-width reference state (tests/golden/net_trained_*.npz) -- well conditioned, so
    two drivers of the same kernels must agree to round-off.  Without: a small default-initialised model."""
    from mmlf_b200.model.feed_forward import FeedForward
    if golden is not None:
        from test_trained_fixtures import build_state
        state, _ = build_state(golden, variant)
        m = FeedForward(**fx.model_kwargs(variant, False, chs=70))
        m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in state.items()})
        return m.cuda()
    torch.manual_seed(seed)
    m = FeedForward(**fx.model_kwargs(variant, False, chs=chs))
    with torch.no_grad():
        fx.perturb_state(m.state_dict(), seed, wscale=1.0)
        m.out_net[7][0].bias.fill_(0.5)                      # keeps the head's ReLU alive
    return m.cuda()


def _batches(n, B=8, ps=32, K=3):
    out = []
    for k in range(n):
        (h, v, i, d), gt, mask = fx.trained_batch(k)
        out.append(dict(views=[T(a) for a in (h, v, i, d)], gt=T(gt), mask=T(mask),
                        mpi=T(fx.synth_mpi(500 + k, gt, K)), pad=T((np.abs(gt) < 1.5).astype(np.int32))))
    return out


def _eager_loss(loss, m, out, b):
    from mmlf_b200.model import loss as L
    from mmlf_b200.utils import dl
    if loss == 'l1':
        return L.MaskedL1Loss()(out, b['gt'], b['mask'])
    if loss == 'multi_l1':
        return L.MultiMaskedL1Loss()(out, b['mpi'], b['mask'])
    if loss == 'upr':
        return L.ImprovedUncertaintyL1Loss()(out, b['gt'], b['mask'])
    if loss == 'upr_pad':
        return L.ImprovedUncertaintyL1Loss()(out, b['gt'], b['mask'], b['pad'])
    if loss == 'multi_upr':
        return L.ImprovedMultiUncertaintyL1Loss()(out, b['mpi'], b['mask'])
    if loss == 'ce':
        return L.MaskedCrossEntropy()(out, dl.reg_to_class(b['gt'], -3.5, 3.5, m.steps), b['mask'])
    if loss == 'ce_mm':
        return L.MaskedCrossEntropy()(out, dl.mpi_to_weights(b['mpi'], -3.5, 3.5, m.steps), b['mask'])
    raise KeyError(loss)


CASES = [('base', 'l1'), ('base', 'multi_l1'), ('upr', 'upr'), ('upr', 'upr_pad'), ('upr', 'multi_upr'), ('dpp', 'ce'),
         ('dpp', 'ce_mm'), ('upr', 'l1')]


@pytest.mark.parametrize('variant,loss', CASES)
@pytest.mark.parametrize('graph', [True, False])
def test_train_step_equals_the_eager_autograd_path(golden, variant, loss, graph):
    """Same kernels, two drivers: `loss_fn(model(...)).backward(); opt.step()` through autograd vs TrainStep (graph
    captured at the first call, replayed afterwards).  Losses, parameters, Adam moments and BatchNorm statistics agree to
    float round-off after 4 steps on changing batches and a changing learning rate."""
    from mmlf_b200.model import loss as L_  # noqa: F401
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    from mmlf_b200.utils import dl
    batches = _batches(2)
    m_e = _model(variant, golden)
    m_s = copy.deepcopy(m_e)
    opt_e, opt_s = FusedAdam(m_e.parameters(), lr=1e-3), FusedAdam(m_s.parameters(), lr=1e-3)
    name = {'upr_pad': 'upr', 'ce_mm': 'ce'}.get(loss, loss)
    step = TrainStep(m_s, opt_s, name, ce_from_gt=(loss == 'ce'), use_graph=graph)
    m_e.train(), m_s.train()
    lrs = [1e-4, 5e-5, 0.0, 2e-4]
    for it in range(4):
        b = batches[it % 2]
        for o in (opt_e, opt_s):
            o.param_groups[0]['lr'] = lrs[it]
        opt_e.zero_grad()
        le = _eager_loss(loss, m_e, m_e(*b['views']), b)
        le.backward()
        opt_e.step()
        if name == 'ce':
            tgt = b['gt'] if loss == 'ce' else dl.mpi_to_weights(b['mpi'], -3.5, 3.5, m_s.steps)
        else:
            tgt = b['mpi'] if name.startswith('multi') else b['gt']
        ls = step(*b['views'], tgt, b['mask'], b['pad'] if loss == 'upr_pad' else None)
        # same kernels, different atomic-add orders in the statistics: round-off, amplified a little by the perturbed fixture
        assert abs(le.item() - ls.item()) <= 1e-3 * abs(le.item()) + 1e-6, (it, le.item(), ls.item())
    assert opt_e.host_step() == opt_s.host_step() == 4
    if graph:
        assert step.replays == 3 and step.launches_per_step > 100
    for (n, pe), (_, ps_) in zip(m_e.named_parameters(), m_s.named_parameters()):
        # Adam turns a gradient element whose sign is round-off (conv biases in front of a BatchNorm: exactly zero in exact
        # arithmetic) into a full +-lr step, so single elements may differ by 2 * sum(lr); tensors are compared by norm
        assert float((pe - ps_).norm()) <= 2e-3 * float(pe.norm()) + 2 * sum(lrs) * pe.numel() ** 0.5 * (
            1.0 if n.endswith('.2.bias') else 0.02), n
    for (n, be), (_, bs_) in zip(m_e.named_buffers(), m_s.named_buffers()):
        assert torch.allclose(be.float(), bs_.float(), rtol=5e-3, atol=2e-4), n
    se, ss = opt_e.state_dict()['state'], opt_s.state_dict()['state']
    names = [n for n, _ in m_e.named_parameters()]
    for k in se:
        if names[k].endswith('.2.bias') and not names[k].startswith('out_net.7.'):
            continue                                   # zero gradient up to round-off (a BatchNorm follows)
        a, b = se[k]['exp_avg'].double(), ss[k]['exp_avg'].double()
        # the two models drift apart by Adam's +-lr steps on noise-level gradient elements; a weight change of that size
        # re-rolls the bf16 roundings of the gradient path (3 % per tensor against the fp32 reference): measured 2-3 %
        assert float((a - b).norm()) <= 6e-2 * float(a.norm()) + 1e-12, names[k]
        assert float(se[k]['step']) == float(ss[k]['step']) == 4.0


def test_out_of_band_parameter_writes_reach_the_captured_graphs():
    """A parameter changed behind the step's back (an in-place torch op, a load_state_dict) must be seen by the next
    replay of the TRAINING graph (its weight-pack launch is part of the graph) and must invalidate the INFERENCE graph
    (keyed on version counters and addresses)."""
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    b = _batches(1)[0]
    m = _model('upr')
    opt = FusedAdam(m.parameters(), lr=0.0)                    # lr 0: the step itself leaves the weights alone
    step = TrainStep(m, opt, 'upr')
    m.train()
    l0 = step(*b['views'], b['gt'], b['mask']).item()
    l1 = step(*b['views'], b['gt'], b['mask']).item()          # replay
    assert step.replays == 1 and abs(l0 - l1) <= 1e-6 * abs(l0)
    m.eval()
    with torch.no_grad():
        e0 = m(*b['views'])['mean'].clone()                    # captures the inference graph
        e0b = m(*b['views'])['mean'].clone()                   # replays it
    assert torch.equal(e0, e0b) and len(m._graphs) == 1
    with torch.no_grad():
        m.out_net[7][0].weight.mul_(1.5)                       # out-of-band write #1: in-place op on a parameter
        m.in_net_hv[0][0].bias.add_(0.05)
    m.train()
    l2 = step(*b['views'], b['gt'], b['mask']).item()          # replay of the SAME training graph
    assert step.replays == 2
    ref = copy.deepcopy(m)
    ref.use_cuda_graph = False
    from mmlf_b200.model import loss as L
    ref.train()
    for bn, rb in zip(m.buffers(), ref.buffers()):
        assert torch.equal(bn, rb)
    with torch.no_grad():
        m.eval(), ref.eval()
        e1 = m(*b['views'])['mean']
        e1_ref = ref(*b['views'])['mean']                      # un-graphed forward of a fresh copy
    assert not torch.allclose(e0, e1) and torch.allclose(e1, e1_ref, rtol=0, atol=1e-6)
    assert abs(l2 - l0) > 1e-4 * abs(l0), 'the replayed training graph did not see the new weights'
    # training loss of the changed weights == eager loss of a fresh copy (BN statistics moved by the steps: copy first)
    ref2 = copy.deepcopy(m).train()
    m.train()
    l3 = step(*b['views'], b['gt'], b['mask']).item()
    l3_ref = L.ImprovedUncertaintyL1Loss()(ref2(*b['views']), b['gt'], b['mask']).item()
    assert abs(l3 - l3_ref) <= 1e-3 * abs(l3_ref) + 1e-6
    # out-of-band write #2: load_state_dict
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd['out_net.7.0.weight'] *= 0.25
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        e2 = m(*b['views'])['mean']
    assert not torch.allclose(e1, e2)


@pytest.mark.parametrize('variant', ['base', 'upr', 'dpp'])
def test_train_step_follows_the_reference_trajectory(golden, variant):
    """The 20-step Adam trajectory of the reference (tests/golden/net_trained_*.npz) through the captured TrainStep."""
    from test_trained_fixtures import build_state
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    state, g = build_state(golden, variant)
    m = FeedForward(**fx.model_kwargs(variant, False, chs=70))
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in state.items()})
    m = m.cuda().train()
    c = fx.TRAINED
    opt = FusedAdam(m.parameters(), lr=c['traj_lr'])
    step = TrainStep(m, opt, {'base': 'l1', 'upr': 'upr', 'dpp': 'ce'}[variant], ce_from_gt=True)
    batches = []
    for k in range(c['n_batches']):
        (h, v, i, d), gt, mask = fx.trained_batch(k)
        batches.append(([T(a) for a in (h, v, i, d)], T(gt), T(mask)))
    traj = []
    for s in range(c['traj_steps']):
        views, gt_t, mask_t = batches[s % c['n_batches']]
        traj.append(step(*views, gt_t, mask_t).item())
    ref = g['traj/loss']
    rel = np.abs(np.array(traj) - ref) / np.maximum(np.abs(ref), 0.5 * np.abs(ref).max())
    assert step.replays == c['traj_steps'] - 1
    assert rel.max() <= 0.03, (traj, ref.tolist())


@pytest.mark.parametrize('over', [dict(model_cross=True), dict(model_no_batchnorm=True),
                                  dict(model_in_blocks=2, model_out_blocks=3, model_views=5)])
def test_train_step_on_other_topologies(over):
    """--model_cross (two stacks), --model_no_batchnorm and a shallower network through the captured step: three steps on
    one batch, the second and third replayed; same losses as the eager autograd path on a twin model."""
    from mmlf_b200.model import loss as L
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    n = over.get('model_views', 9)
    torch.manual_seed(5)
    m_e = FeedForward(**fx.model_kwargs('upr', over.get('model_cross', False), chs=16, **{k: v for k, v in over.items() if k != 'model_cross'}))
    with torch.no_grad():
        m_e.out_net[len(m_e.out_net) - 1][0].bias.fill_(0.5)
    m_e = m_e.cuda()
    m_s = copy.deepcopy(m_e)
    h, v, i, d, gt = fx.synth_batch(310, 4, 24, 24, n=n)
    views = [T(a) for a in (h, v, i, d)]
    gt_t, mask_t = T(gt), T(fx.synth_mask(311, 4, 24, 24))
    opt_e, opt_s = FusedAdam(m_e.parameters(), lr=1e-4), FusedAdam(m_s.parameters(), lr=1e-4)
    step = TrainStep(m_s, opt_s, 'upr')
    m_e.train(), m_s.train()
    for it in range(3):
        opt_e.zero_grad()
        le = L.ImprovedUncertaintyL1Loss()(m_e(*views), gt_t, mask_t)
        le.backward()
        opt_e.step()
        ls = step(*views, gt_t, mask_t)
        assert np.isfinite(ls.item()) and abs(le.item() - ls.item()) <= 2e-3 * abs(le.item()) + 1e-5, (it, le.item(), ls.item())
    assert step.replays == 2
