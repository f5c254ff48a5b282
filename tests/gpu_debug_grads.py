"""Debug helper (run on the GPU box): per-parameter gradient errors of one tiny fixture."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _fixtures as fx, oracle
from oracle import losses as olosses
from mmlf_b200.model.feed_forward import FeedForward
from mmlf_b200.model import loss as L
name = sys.argv[1] if len(sys.argv) > 1 else 'net_tiny_base_full'
prec = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
state = {k[6:]: g[k] for k in g.files if k.startswith('state/')}
kw = fx.model_kwargs('base', 'cross' in name, chs=8)
m = FeedForward(**kw); m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in state.items()}); m = m.cuda()
m.precision = prec
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
h, v, i, d, gt = fx.synth_batch(21, 2, 20, 20); mask = fx.synth_mask(22, 2, 20, 20)
m.train()
out = m(T(h), T(v), T(i), T(d)); loss = L.MaskedL1Loss()(out, T(gt), T(mask)); loss.backward()
emu = oracle.FeedForwardOracle(state, quant=prec, model_cross='cross' in name); emu.training = True
e = emu.forward(h, v, i, d, keep_tape=True)
ev, eg = olosses.masked_l1({'mean': e['mean']}, gt, mask)
egr = emu.backward(eg['mean'][:, None])
print('loss', loss.item(), ev, float(g['train/loss']))
for pname, p in m.named_parameters():
    got = p.grad.cpu().numpy(); ref = g['grad/' + pname]; em = egr[pname]
    print(f'{pname:28s} max|ref| {np.abs(ref).max():.3e} err_ref {np.abs(got-ref).max():.3e} err_emu {np.abs(got-em).max():.3e} emu_vs_ref {np.abs(em-ref).max():.3e}')
