"""GPU twins of the distribution metrics of the reference's validation script (/root/reference/mmlf/validate/cli.py):
``laplace_to_discrete`` / ``lmm_to_discrete`` (:91-118), ``mean_to_discrete`` (:121-137), ``multimodal_mask`` (:165-171),
``kl_divergence`` (:174-187) and ``nll_discrete`` (:51-70) on CUDA tensors, in float64 like the numpy originals -- the
(70 | 108, 512, 512) posteriors never leave the GPU.  ``kl_divergence`` / ``nll_discrete`` normalise their arguments IN
PLACE exactly as the reference does (validate.main calls kl_divergence three times on the same arrays and the later
calls see the earlier normalisation)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._lib import call


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def lmm_to_discrete(n_bins, x_min, x_max, means, logvars):
    """means / logvars: (K, B, H, W) float32 CUDA tensors -> (B, n_bins, H, W) float64."""
    _lib.require_device()
    means, logvars = means.contiguous().float(), logvars.contiguous().float()
    K, B, H, W = means.shape
    out = torch.empty((B, n_bins, H, W), dtype=torch.float64, device=means.device)
    call('mmlf_lmm_to_discrete', _p(means), _p(logvars), K, B, H * W, int(n_bins), float(x_min), float(x_max), _p(out), _st())
    return out


def laplace_to_discrete(n_bins, x_min, x_max, mean, logvar):
    return lmm_to_discrete(n_bins, x_min, x_max, mean.unsqueeze(0), logvar.unsqueeze(0))


def mean_to_discrete(n_bins, x_min, x_max, mean):
    step = (x_max - x_min) / n_bins
    bins = torch.from_numpy(np.linspace(x_min, x_max, n_bins)).to(mean.device).view(1, -1, 1, 1)
    return (torch.abs(bins - mean.unsqueeze(1).double()) < step / 2.0).double()


def multimodal_mask(mpi, threshhold=0.3):
    return ((mpi[:, :, 3] > threshhold).sum(1) > 1).double()


def _metric(name, a, b, mask):
    _lib.require_device()
    for t in (a, b):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise RuntimeError('mmlf_b200: distributions must be contiguous float64 CUDA tensors (normalised in place)')
    B, S, H, W = a.shape
    sums = torch.zeros(2, dtype=torch.float64, device=a.device)
    m = None
    if mask is not None:
        m = torch.as_tensor(mask, dtype=torch.float64, device=a.device).contiguous()
    call(name, _p(a), _p(b), S, B, H * W, _p(m), C.c_void_p(0), _p(sums), _st())
    return (sums[0] / sums[1]).item()


def kl_divergence(dist, dist_gt, mask=None):
    """(B, S, H, W) float64, both normalised in place; masked mean of sum_c gt log(gt / dist)."""
    return _metric('mmlf_kl_divergence', dist, dist_gt, mask)


def nll_discrete(weights, posterior, vmin=None, vmax=None, mask=None):
    return _metric('mmlf_nll_discrete', weights, posterior, mask)
