"""Oracle (test infrastructure): the ESE shift ensemble.

Restates /root/reference/mmlf/model/ensamble.py:40-118.
"""
import numpy as np

from .lf import shift, ese_shift_values
from .net import laplacian, np_linspace_f32


def ensemble_reduce(means, logvars, disp_min, disp_max):
    """ensamble.py:78-101.  means/logvars: (K,B,H,W) float32, ``means`` already
    including ``+ shift_disp``.  Returns mean, logvar (B,H,W) of the member with
    minimal logvar (first minimum on ties, like torch.min) and the Laplace
    mixture posterior (B,K,H,W) on K = len(means) inclusive linspace points."""
    K = means.shape[0]
    idx = np.argmin(logvars, 0)[None]
    mean = np.take_along_axis(means, idx, 0)[0]
    logvar = np.take_along_axis(logvars, idx, 0)[0]
    disp = np_linspace_f32(disp_min, disp_max, K)
    post = np.zeros((means.shape[1], K) + means.shape[2:], np.float32)
    for i in range(K):
        post += laplacian(disp, means[i], np.exp(logvars[i]))
    post /= np.float32(float(K))
    return mean, logvar, post


def ensemble_forward(net, h, v, i, d, disp_min, disp_max, disp_step):
    """ensamble.py:61-76: one UPR forward per ``np.arange`` shift value on shifted
    clones of the four stacks; ``net`` is a FeedForwardOracle in eval mode."""
    means, logvars = [], []
    for s in ese_shift_values(disp_min, disp_max, disp_step):
        hh, vv, ii, dd = shift((h, v, i, d), s)
        out = net.forward(hh, vv, ii, dd)
        means.append((out['mean'] + np.float32(s)).astype(np.float32))
        logvars.append(out['logvar'])
    means, logvars = np.stack(means), np.stack(logvars)
    mean, logvar, post = ensemble_reduce(means, logvars, disp_min, disp_max)
    return {'mean': mean, 'logvar': logvar, 'means': means, 'logvars': logvars, 'posterior': post}
