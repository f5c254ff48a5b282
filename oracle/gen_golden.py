"""Generate tests/golden/* by running the UNMODIFIED reference (titus-leistner/mmlf)
on CPU.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference holds no golden vectors of its own; these files
are what pins the oracle (tests/test_oracle_golden.py) and, through it and
directly, the CUDA path.  Inputs come from tests/_fixtures.py (seeded), so the
fixtures store mostly outputs.
"""
import inspect
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.dont_write_bytecode = True

import _fixtures as fx  # noqa: E402
from mmlf.data import hci4d  # noqa: E402
from mmlf.model import loss as rloss  # noqa: E402
from mmlf.model.ensamble import Ensamble  # noqa: E402
from mmlf.model.feed_forward import FeedForward  # noqa: E402
from mmlf.utils import dl  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)
T = torch.from_numpy


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print('wrote', name, sum(np.asarray(a).nbytes for a in arrs.values()) // 1024, 'KiB raw')


# ---------------------------------------------------------------------------- a1
def gen_indices():
    src = inspect.getsource(hci4d.HCI4D.load_scene).splitlines()
    start = next(i for i, l in enumerate(src) if 'w, h = self.nviews' in l)
    end = next(i for i, l in enumerate(src) if 'dds = ' in l)
    body = '\n'.join(l.strip() for l in src[start:end + 1])
    res = {}
    for n in (9, 7, 5):
        class _S:
            nviews = (n, n)
        env = {'self': _S}
        exec(body, env)  # the reference's own index lines (hci4d.py:142-149)
        res.update({f'us{n}': env['us'], f'vs{n}': env['vs'], f'ids{n}': env['ids'], f'dds{n}': env['dds']})
    save('indices.npz', **{k: np.array(v) for k, v in res.items()})


# ---------------------------------------------------------------------------- a2
def gen_shift():
    rng = np.random.RandomState(7)
    H, W = 12, 12
    base = [rng.uniform(0, 1, (9, 3, H, W)).astype(np.float32) for _ in range(4)]
    gt = rng.uniform(-2, 2, (H, W)).astype(np.float32)
    mpi = rng.uniform(-2, 2, (2, 5, H, W)).astype(np.float64)
    arange = np.arange(-3.5, 3.5, 0.1)
    disps = [2.5, -1.3, 0.0, 1.0, -2.0, 0.49999, float(arange[35]), float(arange[3]), float(arange[69]), 7.25, -3.0]
    out = {'disps': np.array(disps), 'gt': gt, 'mpi': mpi}
    for k, b in enumerate(base):
        out[f'in{k}'] = b
    for j, disp in enumerate(disps):
        data = [b.copy() for b in base] + [np.zeros(1), gt.copy(), mpi.copy()]
        res = hci4d.Shift(float(disp))(tuple(data))
        for k in range(4):
            out[f'out{j}_{k}'] = res[k]
        out[f'gt{j}'] = res[5]
        out[f'mpi{j}'] = res[6]
        # torch branch with a batch dimension, as Ensamble uses it (ensamble.py:63-70)
        tdata = tuple(T(np.stack([b, b[::-1].copy()])) for b in base)
        tres = hci4d.Shift(float(disp))(tdata)
        for k in range(4):
            assert np.array_equal(tres[k][0].numpy(), res[k]), 'numpy and torch Shift differ'
            out[f'tout{j}_{k}'] = tres[k][1].numpy()
    # non-square image for the h/v stacks only is not reachable (Shift always touches 4 stacks)
    save('shift.npz', **out)


# ---------------------------------------------------------------------------- bins
def gen_bins():
    out = {}
    for n in (54, 70, 108):
        out[f'torch{n}'] = torch.linspace(-3.5, 3.5, n).numpy()
        out[f'numpy{n}'] = torch.zeros(n).copy_(T(np.linspace(-3.5, 3.5, n))).numpy()
    out['torch_odd'] = torch.linspace(-1.25, 2.0, 37).numpy()
    rng = np.random.RandomState(3)
    gt = rng.uniform(-3.7, 3.7, (2, 9, 11)).astype(np.float32)
    mpi = fx.synth_mpi(5, gt)
    out['gt'] = gt
    out['mpi'] = mpi
    for n in (54, 108):
        out[f'reg_to_class{n}'] = dl.reg_to_class(T(gt), -3.5, 3.5, n).numpy()
        out[f'mpi_to_weights{n}'] = dl.mpi_to_weights(T(mpi), -3.5, 3.5, n).numpy()
        oh = torch.zeros(2, n, 9, 11)
        oh.scatter_(1, torch.randint(0, n, (2, 1, 9, 11), generator=torch.Generator().manual_seed(1)), 1.0)
        out[f'onehot{n}'] = oh.numpy()
        out[f'class_to_reg{n}'] = dl.class_to_reg(oh, -3.5, 3.5, n).numpy()
    for m in (0, 3, 11):
        out[f'margin{m}'] = rloss.create_mask_margin((2, 30, 26), m).numpy()
    save('bins.npz', **out)


# ---------------------------------------------------------------------------- net
def _loss_for(variant, multimodal):
    if variant == 'upr':
        return rloss.ImprovedMultiUncertaintyL1Loss() if multimodal else rloss.ImprovedUncertaintyL1Loss()
    if variant == 'dpp':
        return rloss.MaskedCrossEntropy()
    return rloss.MultiMaskedL1Loss() if multimodal else rloss.MaskedL1Loss()


def _calibrate_bn(model, kw, seed, H, W):
    """Eval-mode fixtures need running statistics that match the (scaled) weights,
    otherwise activations explode through 22 convs.  Take them from one train-mode
    pass with momentum 1 on a calibration input, then detune them a little so that
    eval-mode and train-mode normalisation differ.  The resulting running stats
    are stored in the fixture ('state/*running*')."""
    if kw['model_no_batchnorm']:
        return
    cal = FeedForward(**dict(kw, model_batchnorm_momentum=1.0))
    cal.load_state_dict(model.state_dict())
    cal.train()
    h, v, i, d, _ = fx.synth_batch(seed + 100, 2, H, W, n=kw['model_views'])
    cal(T(h), T(v), T(i), T(d))
    rng = np.random.RandomState(seed + 1)
    sd = model.state_dict()
    for k, t in cal.state_dict().items():
        if k.endswith('running_var'):
            sd[k].copy_(t * T(rng.uniform(0.8, 1.25, tuple(t.shape)).astype(np.float32)))
        elif k.endswith('running_mean'):
            sd[k].copy_(t + 0.1 * T(rng.uniform(-1, 1, tuple(t.shape)).astype(np.float32)))


def _run_net(kw, state_seed, B, H, W, in_seed, store_state, multimodal=False, sample_stride=None, wscale=2.0):
    torch.manual_seed(0)
    model = FeedForward(**kw)
    sd = model.state_dict()
    with torch.no_grad():
        fx.perturb_state(sd, state_seed, wscale=wscale)
        _calibrate_bn(model, kw, state_seed, H, W)
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if store_state or 'running' in k:
            out['state/' + k] = v.numpy().copy()
    h, v, i, d, gt = fx.synth_batch(in_seed, B, H, W, n=kw['model_views'])
    mask = fx.synth_mask(in_seed + 1, B, H, W)
    mpi = fx.synth_mpi(in_seed + 2, gt)
    args = [T(h), T(v), T(i), T(d)]
    variant = 'upr' if kw['model_uncert'] else ('dpp' if kw['model_discrete'] else 'base')
    # ---- eval forward
    model.eval()
    with torch.no_grad():
        o = model(*[a.clone() for a in args])
    for k, t in o.items():
        if t is not None:
            out['eval/' + k] = t.numpy()
    # ---- train forward + loss + backward
    model.train()
    o = model(*[a.clone() for a in args])
    if variant == 'dpp':
        if multimodal:
            target = dl.mpi_to_weights(T(mpi), kw['val_disp_min'], kw['val_disp_max'], model.steps)
        else:
            target = dl.reg_to_class(T(gt), kw['val_disp_min'], kw['val_disp_max'], model.steps)
    else:
        target = T(mpi) if multimodal else T(gt)
    lossv = _loss_for(variant, multimodal)(o, target, T(mask))
    lossv.backward()
    out['train/loss'] = np.array(lossv.item(), np.float64)
    for k, t in o.items():
        if t is not None and k in ('mean', 'logvar', 'scores'):
            out['train/' + k] = t.detach().numpy()
    for name, p in model.named_parameters():
        g = p.grad.numpy()
        if sample_stride and g.size > 4096:
            g = g.reshape(-1)[::sample_stride].copy()
        out['grad/' + name] = g
    for k, vv in model.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            out['after/' + k] = vv.numpy().copy()
    return out


def gen_net():
    for variant in ('base', 'upr', 'dpp'):
        for cross in (False, True):
            for mm in (False, True):
                if cross and mm:
                    continue
                kw = fx.model_kwargs(variant, cross, chs=8)
                res = _run_net(kw, 11, 2, 20, 20, 21, store_state=True, multimodal=mm)
                save(f'net_tiny_{variant}_{"cross" if cross else "full"}{"_mm" if mm else ""}.npz', **res)
    # full-width models: parameters are re-created from the seed in the tests (not stored)
    for variant in ('base', 'upr', 'dpp'):
        kw = fx.model_kwargs(variant, False, chs=70)
        res = _run_net(kw, 13, 2, 16, 16, 31, store_state=False, sample_stride=97)
        save(f'net_full_{variant}.npz', **res)
    kw = fx.model_kwargs('base', True, chs=70)
    save('net_full_base_cross.npz', **_run_net(kw, 13, 1, 16, 16, 33, store_state=False, sample_stride=97))
    # another topology: --model_in_blocks 2 --model_out_blocks 4 --model_views 7 (84 bins)
    kw = fx.model_kwargs('upr', False, chs=8, model_in_blocks=2, model_out_blocks=4, model_views=7)
    save('net_tiny_upr_topo247.npz', **_run_net(kw, 11, 2, 20, 20, 21, store_state=True))
    # odd --model_ksize (symmetric padding k // 2 for both convs of a block, feed_forward.py:86-88): oracle-only fixture, the
    # CUDA path implements the published ksize = 2
    kw = fx.model_kwargs('base', False, chs=8, model_ksize=3)
    save('net_tiny_base_k3.npz', **_run_net(kw, 11, 2, 20, 20, 21, store_state=True, wscale=1.0))
    # no-batchnorm topology (state_dict index 3 vanishes, feed_forward.py:132-135)
    kw = fx.model_kwargs('base', False, chs=8, model_no_batchnorm=True)
    # without BN the scale of the activations is set by the weights alone: x2.8 keeps the output from collapsing
    save('net_tiny_base_nobn.npz', **_run_net(kw, 11, 2, 20, 20, 21, store_state=True, wscale=2.8))


def gen_unet():
    """--model_unet (feed_forward.py:99-100, 189-204; unet.py): UPR model with the U-Net out-net.  The 31 M parameters are
    not stored: names + shapes are, and fixtures.synth_state regenerates the values from the seed."""
    kw = fx.model_kwargs('upr', False, chs=8, model_unet=True)
    torch.manual_seed(0)
    model = FeedForward(**kw)
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    state = fx.synth_state(shapes, 17)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    B, H, W = 2, 32, 32
    h, v, i, d, gt = fx.synth_batch(51, B, H, W)
    mask = fx.synth_mask(52, B, H, W)
    out = {'names': np.array([k for k, _ in shapes]), 'shapes': np.array([','.join(map(str, s)) for _, s in shapes])}
    model.eval()
    with torch.no_grad():
        o = model(T(h), T(v), T(i), T(d))
    out['eval/mean'], out['eval/logvar'] = o['mean'].numpy(), o['logvar'].numpy()
    model.train()
    o = model(T(h), T(v), T(i), T(d))
    lossv = _loss_for('upr', False)(o, T(gt), T(mask))
    lossv.backward()
    out['train/loss'] = np.array(lossv.item(), np.float64)
    out['train/mean'], out['train/logvar'] = o['mean'].detach().numpy(), o['logvar'].detach().numpy()
    for name, p in model.named_parameters():
        g = p.grad.numpy()
        out['grad/' + name] = g.reshape(-1)[::97].copy() if g.size > 4096 else g
    for k, vv in model.state_dict().items():
        if 'running' in k and 'out_net' in k and vv.numel() <= 128:
            out['after/' + k] = vv.numpy().copy()
    save('net_unet_upr.npz', **out)


def gen_evalmode():
    """--train_eval_mode (train/cli.py:227-230): the training step with the model in eval() mode, i.e. gradients through
    BatchNorm layers that normalise with their running statistics (which must not change)."""
    kw = fx.model_kwargs('base', False, chs=8)
    torch.manual_seed(0)
    model = FeedForward(**kw)
    with torch.no_grad():
        fx.perturb_state(model.state_dict(), 11, wscale=2.0)
        _calibrate_bn(model, kw, 11, 20, 20)
    out = {'state/' + k: v.numpy().copy() for k, v in model.state_dict().items()}
    h, v, i, d, gt = fx.synth_batch(23, 2, 20, 20)
    mask = fx.synth_mask(24, 2, 20, 20)
    model.eval()
    o = model(T(h), T(v), T(i), T(d))
    lossv = _loss_for('base', False)(o, T(gt), T(mask))
    lossv.backward()
    out['loss'] = np.array(lossv.item(), np.float64)
    out['mean'] = o['mean'].detach().numpy()
    for name, p in model.named_parameters():
        out['grad/' + name] = p.grad.numpy()
    for k, vv in model.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            assert np.array_equal(vv.numpy(), out['state/' + k])          # eval mode leaves the statistics alone
    save('net_tiny_base_evalmode.npz', **out)


# ---------------------------------------------------------------------------- trained-like full-width fixtures
TRAINED = dict(B=8, ps=32, lr=1e-3, traj_lr=2e-4, trunk_steps=30, head_steps=60, traj_steps=20, n_batches=4, seed0=200,
               grad_stride=13)


def _trained_batch(k):
    """Batch k of the trained-like fixtures: 8 synthetic 32 x 32 light fields + mask + multi-plane target."""
    c = TRAINED
    h, v, i, d, gt = fx.synth_batch(c['seed0'] + k, c['B'], c['ps'], c['ps'])
    mask = fx.synth_mask(c['seed0'] + 50 + k, c['B'], c['ps'], c['ps'], margin=3)
    return [T(h), T(v), T(i), T(d)], gt, mask


def _variant_loss(model, variant, o, gt, mask):
    if variant == 'dpp':
        target = dl.reg_to_class(T(gt), -3.5, 3.5, model.steps)
        return rloss.MaskedCrossEntropy()(o, target, T(mask))
    return _loss_for(variant, False)(o, T(gt), T(mask))


def gen_trained():
    """Well-conditioned, non-degenerate full-width (chs = 70) states: the reference model after real training steps.

    The trunk (in-nets + out-net blocks 0..6) is trained as a BASE model for 30 Adam steps (lr 1e-3, B = 8, 32 px,
    train-mode BatchNorm) from the torch.manual_seed(0) default init; its conv-weight DELTAS are quantised to int8 per
    tensor so that the 4.6 M parameters cost 4.6 MB once (the init is re-created from the seed by the tests).  UPR and
    DPP take the same trunk and train their own head (trunk frozen).  Everything the fixture records -- eval outputs,
    one training step (loss, outputs, gradients), a 20-step Adam loss trajectory -- is computed by the reference FROM THE
    QUANTISED STATE, i.e. from exactly the weights the tests rebuild."""
    c = TRAINED
    batches = [_trained_batch(k) for k in range(c['n_batches'])]
    kw = fx.model_kwargs('base', False, chs=70)
    torch.manual_seed(0)
    model = FeedForward(**kw)
    init = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    trunk_file = os.path.join(OUT, 'net_trained_trunk.npz')
    if os.environ.get('KEEP_TRUNK') == '1' and os.path.exists(trunk_file):
        trunk = dict(np.load(trunk_file))                 # regenerate the per-variant files on the committed trunk
    else:
        opt = torch.optim.Adam(model.parameters(), lr=c['lr'])
        model.train()
        for s in range(c['trunk_steps']):
            args, gt, mask = batches[s % c['n_batches']]
            opt.zero_grad()
            lossv = _variant_loss(model, 'base', model(*args), gt, mask)
            lossv.backward()
            opt.step()
            print('trunk step', s, lossv.item())
        trunk = {}
        for k, v in model.state_dict().items():
            if k.startswith('out_net.7.'):
                continue
            a = v.detach().numpy()
            if a.ndim == 4:                                   # conv weight: int8 delta from the seeded init
                q, scale = fx.quantise_delta(a, init[k])
                trunk['q/' + k], trunk['s/' + k] = q, scale
            else:
                trunk['f/' + k] = a.copy()
        save('net_trained_trunk.npz', **trunk)
    trunk_state = fx.trained_trunk_state(trunk, init)

    for variant in ('base', 'upr', 'dpp'):
        kw = fx.model_kwargs(variant, False, chs=70)
        torch.manual_seed(0)
        model = FeedForward(**kw)
        sd = model.state_dict()
        for k, a in trunk_state.items():
            sd[k].copy_(T(np.array(a)))
        # ---- head: trained with the trunk frozen
        head = [p for n, p in model.named_parameters() if n.startswith('out_net.7.')]
        for n, p in model.named_parameters():
            p.requires_grad_(n.startswith('out_net.7.'))
        opt = torch.optim.Adam(head, lr=c['lr'])
        model.train()
        for s in range(c['head_steps']):
            args, gt, mask = batches[s % c['n_batches']]
            opt.zero_grad()
            lossv = _variant_loss(model, variant, model(*args), gt, mask)
            lossv.backward()
            opt.step()
            if s % 10 == 0:
                print(variant, 'head step', s, lossv.item())
        for p in model.parameters():
            p.requires_grad_(True)
            p.grad = None
        out = {}
        for k, v in model.state_dict().items():
            if k.startswith('out_net.7.') or 'running' in k or 'num_batches' in k:
                out['state/' + k] = v.detach().numpy().copy()
        start_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        # ---- eval forward on batch 0
        args, gt, mask = batches[0]
        model.eval()
        with torch.no_grad():
            o = model(*[a.clone() for a in args])
        for k, t in o.items():
            if t is None:
                continue
            a = t.numpy()
            if a.ndim == 4 and a.shape[1] > 2:            # 108-plane tensors: light fields 0 and 5 only
                a = a[[0, 5]]
            out['eval/' + k] = a.copy()
        # ---- one training step (forward + loss + backward) on batch 0
        model.train()
        o = model(*[a.clone() for a in args])
        lossv = _variant_loss(model, variant, o, gt, mask)
        lossv.backward()
        out['train/loss'] = np.array(lossv.item(), np.float64)
        for k in ('mean', 'logvar', 'scores'):
            if o[k] is not None:
                a = o[k].detach().numpy()
                out['train/' + k] = (a[[0, 5]] if a.ndim == 4 else a).copy()
        for name, p in model.named_parameters():
            g = p.grad.numpy()
            out['grad/' + name] = g.reshape(-1)[::c['grad_stride']].copy() if g.size > 4096 else g.copy()
            out['gnorm/' + name] = np.array(np.sqrt((g.astype(np.float64) ** 2).sum()))
        for k, vv in model.state_dict().items():
            if 'running' in k or 'num_batches' in k:
                out['after/' + k] = vv.numpy().copy()
        # ---- 20-step Adam trajectory from the start state (train/cli.py:243-258: forward, loss, backward, step)
        model.load_state_dict(start_state)
        opt = torch.optim.Adam(model.parameters(), lr=c['traj_lr'])
        traj = []
        for s in range(c['traj_steps']):
            args, gt, mask = batches[s % c['n_batches']]
            opt.zero_grad()
            lossv = _variant_loss(model, variant, model(*args), gt, mask)
            lossv.backward()
            opt.step()
            traj.append(lossv.item())
        out['traj/loss'] = np.array(traj, np.float64)
        with torch.no_grad():                     # training mode (batch statistics): see the test for why not eval()
            o = model(*[a.clone() for a in batches[0][0]])
        key = 'scores' if variant == 'dpp' else 'mean'
        a = o[key].numpy()
        out['traj/final_train_' + key] = (a[[0, 5]] if a.ndim == 4 else a).copy()
        print(variant, 'trajectory', traj[0], '->', traj[-1])
        save(f'net_trained_{variant}.npz', **out)


def gen_randomshift():
    """RandomShift (hci4d.py:993-1028) under random.seed(s): the drawn disparity and the resampled stacks, numpy branch."""
    import random
    rng = np.random.RandomState(31)
    H = W = 16
    base = [rng.uniform(0, 1, (9, 3, H, W)).astype(np.float32) for _ in range(4)]
    gt = rng.uniform(-2, 2, (H, W)).astype(np.float32)
    mpi = rng.uniform(-2, 2, (2, 5, H, W)).astype(np.float64)
    out = {'gt': gt, 'mpi': mpi}
    for k, b in enumerate(base):
        out[f'in{k}'] = b
    cases = [(101, 1.0), (102, 1.0), (103, 2.5), (104, (-0.5, 3.0)), (105, (1.0, 1.0))]
    for j, (seed, rg) in enumerate(cases):
        random.seed(seed)
        data = [b.copy() for b in base] + [np.zeros(1), gt.copy(), mpi.copy()]
        res = hci4d.RandomShift(rg)(tuple(data))
        for k in range(4):
            out[f'out{j}_{k}'] = res[k]
        out[f'gt{j}'], out[f'mpi{j}'] = res[5], res[6]
        out[f'disp{j}'] = np.array(float(gt[0, 0]) - float(res[5][0, 0]))
    out['seeds'] = np.array([s for s, _ in cases])
    out['ranges'] = np.array([(r if isinstance(r, tuple) else (-r, r)) for _, r in cases], np.float64)
    out['is_tuple'] = np.array([isinstance(r, tuple) for _, r in cases])
    save('randomshift.npz', **out)



# ---------------------------------------------------------------------------- losses
def gen_losses():
    rng = np.random.RandomState(17)
    B, H, W, n = 2, 14, 18, 108
    gt = rng.uniform(-2, 2, (B, H, W)).astype(np.float32)
    mean = (gt + rng.normal(0, 0.3, gt.shape)).astype(np.float32)
    logvar = rng.normal(-0.5, 0.7, gt.shape).astype(np.float32)
    scores = rng.normal(0, 1.5, (B, n, H, W)).astype(np.float32)
    mask = fx.synth_mask(4, B, H, W)
    mpi = fx.synth_mpi(6, gt)
    mask_padding = (np.abs(gt) < 1.5).astype(np.int32)
    out = dict(gt=gt, mean=mean, logvar=logvar, scores=scores, mask=mask, mpi=mpi, mask_padding=mask_padding)

    def run(name, fn, target, *extra, keys=('mean',), m=mask):
        o = {'mean': T(mean).requires_grad_(), 'logvar': T(logvar).requires_grad_(),
             'scores': T(scores).requires_grad_()}
        val = fn(o, target, T(m), *extra)
        out[name + '/value'] = np.array(val.item(), np.float64)
        if val.requires_grad:
            val.backward()
            for k in keys:
                out[f'{name}/g_{k}'] = o[k].grad.numpy()

    run('l1', rloss.MaskedL1Loss(), T(gt))
    run('l1_empty', rloss.MaskedL1Loss(), T(gt), m=np.zeros_like(mask))
    run('multi_l1', rloss.MultiMaskedL1Loss(), T(mpi))
    run('mse', rloss.MaskedMSELoss(), T(gt))
    run('badpix', rloss.MaskedBadPix(), T(gt), keys=())
    run('upr', rloss.ImprovedUncertaintyL1Loss(), T(gt), keys=('mean', 'logvar'))
    run('upr_pad', rloss.ImprovedUncertaintyL1Loss(), T(gt), T(mask_padding), keys=('mean', 'logvar'))
    run('multi_upr', rloss.ImprovedMultiUncertaintyL1Loss(), T(mpi), keys=('mean', 'logvar'))
    t1 = dl.reg_to_class(T(gt), -3.5, 3.5, n)
    t2 = dl.mpi_to_weights(T(mpi), -3.5, 3.5, n)
    run('ce', rloss.MaskedCrossEntropy(), t1, keys=('scores',))
    run('ce_mm', rloss.MaskedCrossEntropy(), t2, keys=('scores',))
    save('losses.npz', **out)


# ---------------------------------------------------------------------------- ESE
def gen_ese():
    kw = fx.model_kwargs('upr', False, chs=8)
    torch.manual_seed(0)
    model = FeedForward(**kw)
    sd = model.state_dict()
    with torch.no_grad():
        fx.perturb_state(sd, 19)
        _calibrate_bn(model, kw, 19, 16, 16)
    model.eval()
    h, v, i, d, gt = fx.synth_batch(41, 1, 16, 16)
    out = {'state/' + k: t.numpy().copy() for k, t in model.state_dict().items()}
    for step, tag in ((0.1, 'full'), (1.0, 'coarse')):
        ens = Ensamble(model, -3.5, 3.5, step)
        with torch.no_grad():
            o = ens(T(h), T(v), T(i), T(d))
        for k, t in o.items():
            out[f'{tag}/{k}'] = t.numpy()
    save('ese_tiny.npz', **out)


# ---------------------------------------------------------------------------- Adam
def gen_adam():
    rng = np.random.RandomState(23)
    p0 = rng.normal(0, 1, (5, 7)).astype(np.float32)
    grads = rng.normal(0, 1, (4, 5, 7)).astype(np.float32)
    p = torch.nn.Parameter(T(p0.copy()))
    opt = torch.optim.Adam([p], lr=1e-3)
    out = {'p0': p0, 'grads': grads}
    lrs = [0.0, 1e-3, 1e-3, 2.5e-4]
    for s in range(4):
        for g in opt.param_groups:
            g['lr'] = lrs[s]
        p.grad = T(grads[s].copy())
        opt.step()
        out[f'p{s + 1}'] = p.detach().numpy().copy()
    st = opt.state_dict()['state'][0]
    out['exp_avg'] = st['exp_avg'].numpy()
    out['exp_avg_sq'] = st['exp_avg_sq'].numpy()
    out['lrs'] = np.array(lrs)
    save('adam.npz', **out)


# ---------------------------------------------------------------------------- checkpoint
def gen_checkpoint():
    """A checkpoint.pt written by the reference's own ModelSaver (utils/dl.py:21-74)
    from a random-init reference model + Adam state, as train/cli.py:327-329 does."""
    kw = fx.model_kwargs('upr', False, chs=8)
    hyper = dict(kw, train_lr=1e-3, train_bs=2, train_ps=20, train_shift=2.5, val_ensamble=False,
                 val_disp_step=0.1, model_radius=11, train_loss_multimodal=False)
    torch.manual_seed(5)
    model = FeedForward(**kw)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    h, v, i, d, gt = fx.synth_batch(51, 2, 20, 20)
    o = model(T(h), T(v), T(i), T(d))
    rloss.ImprovedUncertaintyL1Loss()(o, T(gt), T(fx.synth_mask(52, 2, 20, 20))).backward()
    opt.step()
    dl.ModelSaver()(os.path.join(OUT, 'ref_checkpoint_tiny.pt'), torch.nn.DataParallel(model), opt, hyper, None, 1, 0.5)
    model.eval()
    with torch.no_grad():
        o = model(T(h), T(v), T(i), T(d))
    save('ref_checkpoint_tiny_out.npz', mean=o['mean'].numpy(), logvar=o['logvar'].numpy())


# ---------------------------------------------------------------------------- f2
def gen_texture():
    """create_mask_texture (hci4d.py:38-69) on synthetic centre views with flat and textured regions; the windowed
    mean-L1 map is stored too (same torch ops as the reference, before the threshold) so that the parity tests can
    exclude pixels whose mean sits within float round-off of the threshold."""
    rng = np.random.RandomState(17)
    res = {}
    for tag, (b, H, W, ws, thr) in {'a': (1, 48, 64, 23, 0.02), 'b': (2, 40, 40, 7, 0.05)}.items():
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
        c = np.zeros((b, 3, H, W), np.float32)
        for bi in range(b):
            for ch in range(3):
                tex = 0.5 + 0.3 * np.sin(0.9 * xx + ch) * np.cos(0.7 * yy + bi)
                amp = np.clip((xx - W * 0.3) / (W * 0.5), 0, 1)          # flat on the left, textured on the right
                c[bi, ch] = 0.4 + amp * (tex - 0.4) + rng.normal(0, 0.004, (H, W))
        ct = T(c)
        mask = hci4d.create_mask_texture(ct, ws, thr)
        unf = torch.nn.functional.unfold(ct, kernel_size=ws, padding=ws // 2).view(b, 3, -1, H, W)
        mae = torch.abs(unf - ct.unsqueeze(2)).mean((1, 2))
        res.update({f'{tag}/center': c, f'{tag}/mask': mask.numpy().astype(np.int32), f'{tag}/mae': mae.numpy(),
                    f'{tag}/wsize': np.array(ws), f'{tag}/threshold': np.array(thr)})
    save('texture_mask.npz', **res)


# ---------------------------------------------------------------------------- f1
def gen_augment():
    """The reference's own augmentation chain (train/cli.py:78-87) under random.seed(s) on a small synthetic 9-tuple;
    inputs once, outputs + the seed per case.  The parameters are NOT stored: the tests re-derive them by replaying the
    draw order (oracle.augment.draw_params) from the same seed."""
    import random
    from torchvision import transforms
    rng = np.random.RandomState(23)
    n, H, W, ps = 9, 64, 64, 8
    stacks = [rng.uniform(0, 1, (n, 3, H, W)).astype(np.float32) for _ in range(4)]
    center = stacks[1][n // 2].copy()
    gt = rng.uniform(-2, 2, (H, W)).astype(np.float32)
    mpi = np.zeros((1, 5, H, W))
    mpi[0, :3], mpi[0, 3], mpi[0, 4] = center, 1.0, gt
    mask = (rng.uniform(size=(H, W)) > 0.3).astype(np.int64)
    index = np.atleast_1d(3)
    res = {'h': stacks[0], 'v': stacks[1], 'i': stacks[2], 'd': stacks[3], 'center': center, 'gt': gt, 'mpi': mpi,
           'mask': mask, 'ps': np.array(ps), 'max_factor': np.array(2)}
    chain = transforms.Compose([hci4d.RandomDownSampling(2), hci4d.RandomShift(1.0), hci4d.RandomCrop(ps + 2 * 4 * 2),
                                hci4d.CenterCrop(ps), hci4d.RandomRotate(), hci4d.RedistColor(), hci4d.Brightness(),
                                hci4d.Contrast()])
    seeds = list(range(100, 112))
    for s in seeds:
        random.seed(s)
        data = tuple(a.copy() for a in (*stacks, center, gt, mpi, mask, index))
        out = chain(data)
        for k, name in enumerate(('h', 'v', 'i', 'd', 'center', 'gt', 'mpi', 'mask')):
            res[f'{s}/{name}'] = np.ascontiguousarray(out[k])
    res['seeds'] = np.array(seeds)
    save('augment.npz', **res)


# ---------------------------------------------------------------------------- f3
def gen_metrics():
    """The distribution metrics of validate/cli.py on small random inputs, incl. the three consecutive kl_divergence
    calls of validate.main (:323-325) that see each other's in-place normalisation."""
    import contextlib
    import io
    from mmlf.validate import cli as vcli
    rng = np.random.RandomState(29)
    K, B, H, W, S = 5, 1, 6, 7, 108
    means = rng.uniform(-3, 3, (K, B, H, W)).astype(np.float32)
    logvars = rng.normal(-1.0, 0.7, (K, B, H, W)).astype(np.float32)
    mpi = rng.uniform(0, 1, (B, 3, 5, H, W))
    mpi[:, :, 4] = rng.uniform(-3, 3, (B, 3, H, W))
    dist_gt = rng.uniform(0, 1, (B, S, H, W)) * (rng.uniform(size=(B, S, H, W)) > 0.9)
    res = {'means': means, 'logvars': logvars, 'mpi': mpi, 'dist_gt': dist_gt}
    with contextlib.redirect_stdout(io.StringIO()):
        res['laplace'] = vcli.laplace_to_discrete(S, -3.5, 3.5, means[0], logvars[0])
        res['lmm'] = vcli.lmm_to_discrete(S, -3.5, 3.5, means, logvars)
        res['mean_disc'] = vcli.mean_to_discrete(S, -3.5, 3.5, means[0])
        mask = vcli.multimodal_mask(mpi)
        res['mm_mask'] = mask
        d, g = res['lmm'].copy(), dist_gt.copy()
        res['kld'] = np.array([vcli.kl_divergence(d, g), vcli.kl_divergence(d, g, mask), vcli.kl_divergence(d, g, 1.0 - mask)])
        res['kld_dist_after'], res['kld_gt_after'] = d, g
        w, p = dist_gt.copy(), res['lmm'].copy()
        res['nll'] = np.array(vcli.nll_discrete(w, p, -3.5, 3.5, None))
    save('metrics.npz', **res)


# ---------------------------------------------------------------------------- (b) boundary: the CLI surfaces
def gen_cli():
    """Option names, defaults, flags and types of the reference's two click commands (train/cli.py:17-59,
    validate/cli.py:190-208), by introspection."""
    import json
    from mmlf.train import cli as tcli
    from mmlf.validate import cli as vcli
    res = {}
    for name, cmd in (('train', tcli.main), ('validate', vcli.main)):
        res[name] = [{'name': p.name, 'opts': list(p.opts), 'kind': type(p).__name__, 'default': p.default,
                      'is_flag': bool(getattr(p, 'is_flag', False)), 'type': p.type.name} for p in cmd.params]
    with open(os.path.join(OUT, 'cli_options.json'), 'w') as f:
        json.dump(res, f, indent=1, default=str)
    print('wrote cli_options.json', {k: len(v) for k, v in res.items()})


if __name__ == '__main__':
    which = sys.argv[1:] or ['indices', 'shift', 'bins', 'net', 'losses', 'ese', 'adam', 'checkpoint', 'texture', 'augment', 'metrics', 'evalmode', 'unet', 'cli', 'trained', 'randomshift']
    for w in which:
        globals()['gen_' + w]()
