"""Tensor-level wrappers of the C-ABI kernels, registered as torch custom ops (``torch.ops.mmlf.*``).

Each op checks that its tensors are contiguous CUDA tensors of the expected dtype and enqueues the kernel on the
current stream.  There is no fallback implementation: on a CPU tensor the ops raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f'mmlf_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)')
    if t.dtype != dtype:
        raise RuntimeError(f'mmlf_b200: `{name}` must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise RuntimeError(f'mmlf_b200: `{name}` must be contiguous')
    return t


# ----------------------------------------------------------------------------------------- bins (host tables)
def torch_bins(start, stop, steps, device):
    """``torch.linspace`` table as the reference builds it on the CPU (utils/dl.py:126,151,177)."""
    return torch.linspace(start, stop, steps).to(device)


def numpy_bins(start, stop, steps, device):
    """``np.linspace`` cast to f32 (feed_forward.py:287-288, 298-299; ensamble.py:91-92)."""
    return torch.from_numpy(np.linspace(start, stop, steps).astype(np.float32)).to(device)


# ----------------------------------------------------------------------------------------- light field
@torch.library.custom_op('mmlf::lf_extract_u8', mutates_args=())
def lf_extract_u8(views: torch.Tensor, n: int) -> list[torch.Tensor]:
    """views (n*n, H, W, 3) uint8 -> [h, v, i, d (n, 3, H, W) f32, center (3, H, W)]  (hci4d.py:142-193)."""
    _lib.require_device()
    _chk(views, torch.uint8, 'views')
    H, W = views.shape[1], views.shape[2]
    outs = [torch.empty((n, 3, H, W), dtype=torch.float32, device=views.device) for _ in range(4)]
    center = torch.empty((3, H, W), dtype=torch.float32, device=views.device)
    call('mmlf_lf_extract_u8', _p(views), n, H, W, *[_p(o) for o in outs], _p(center), _st())
    return outs + [center]


@lf_extract_u8.register_fake
def _(views, n):
    H, W = views.shape[1], views.shape[2]
    return [views.new_empty((n, 3, H, W), dtype=torch.float32) for _ in range(4)] + \
        [views.new_empty((3, H, W), dtype=torch.float32)]


@torch.library.custom_op('mmlf::lf_shift', mutates_args=())
def lf_shift(h: torch.Tensor, v: torch.Tensor, i: torch.Tensor, d: torch.Tensor, disp: float) -> list[torch.Tensor]:
    """Out-of-place Shift of the four stacks (..., n, 3, H, W) f32 (hci4d.py:907-981)."""
    _lib.require_device()
    for t, nm in ((h, 'h_views'), (v, 'v_views'), (i, 'i_views'), (d, 'd_views')):
        _chk(t, torch.float32, nm)
    n, H, W = h.shape[-4], h.shape[-2], h.shape[-1]
    batch = h.numel() // (n * 3 * H * W)
    outs = [torch.empty_like(t) for t in (h, v, i, d)]
    call('mmlf_lf_shift', _p(h), _p(v), _p(i), _p(d), *[_p(o) for o in outs], batch, n, H, W, float(disp), _st())
    return outs


@lf_shift.register_fake
def _(h, v, i, d, disp):
    return [torch.empty_like(t) for t in (h, v, i, d)]


def texture_mask(center, wsize, threshold, want_mae=False):
    """center (B, 3, H, W) f32 -> int32 mask (B, H, W) [+ the windowed mean-L1 map]  (hci4d.py:38-69)."""
    _lib.require_device()
    center = _chk(center.contiguous(), torch.float32, 'center')
    B, c3, H, W = center.shape
    if c3 != 3:
        raise RuntimeError('mmlf_b200: `center` must have 3 colour planes')
    mask = torch.empty((B, H, W), dtype=torch.int32, device=center.device)
    mae = torch.empty((B, H, W), dtype=torch.float32, device=center.device) if want_mae else None
    call('mmlf_texture_mask', _p(center), B, H, W, int(wsize), float(threshold), _p(mask), _p(mae), _st())
    return (mask, mae) if want_mae else mask


# ----------------------------------------------------------------------------------------- heads
@torch.library.custom_op('mmlf::upr_posterior', mutates_args=())
def upr_posterior(mean: torch.Tensor, logvar: torch.Tensor, bins: torch.Tensor) -> torch.Tensor:
    _lib.require_device()
    mean, logvar = _chk(mean.contiguous(), torch.float32, 'mean'), _chk(logvar.contiguous(), torch.float32, 'logvar')
    B, H, W = mean.shape
    steps = bins.numel()
    post = torch.empty((B, steps, H, W), dtype=torch.float32, device=mean.device)
    call('mmlf_upr_posterior', _p(mean), _p(logvar), _p(bins), steps, B, H * W, _p(post), _st())
    return post


@upr_posterior.register_fake
def _(mean, logvar, bins):
    B, H, W = mean.shape
    return mean.new_empty((B, bins.numel(), H, W))


@torch.library.custom_op('mmlf::dpp_head', mutates_args=())
def dpp_head(scores: torch.Tensor, bins_t: torch.Tensor, bins_n: torch.Tensor) -> list[torch.Tensor]:
    """scores (B, S, H, W) -> [one_hot, posterior, mean, logvar]  (feed_forward.py:276-290)."""
    _lib.require_device()
    _chk(scores, torch.float32, 'scores')
    B, S, H, W = scores.shape
    one_hot = torch.empty_like(scores)
    post = torch.empty_like(scores)
    mean = torch.empty((B, H, W), dtype=torch.float32, device=scores.device)
    logvar = torch.empty_like(mean)
    call('mmlf_dpp_head', _p(scores), _p(bins_t), _p(bins_n), S, B, H * W, _p(one_hot), _p(post), _p(mean),
         _p(logvar), _st())
    return [one_hot, post, mean, logvar]


@dpp_head.register_fake
def _(scores, bins_t, bins_n):
    B, S, H, W = scores.shape
    return [torch.empty_like(scores), torch.empty_like(scores), scores.new_empty((B, H, W)), scores.new_empty((B, H, W))]


@torch.library.custom_op('mmlf::reg_to_class', mutates_args=())
def reg_to_class_op(gt: torch.Tensor, bins_t: torch.Tensor, half_step: float) -> torch.Tensor:
    _lib.require_device()
    _chk(gt, torch.float32, 'gt')
    B, H, W = gt.shape
    S = bins_t.numel()
    out = torch.empty((B, S, H, W), dtype=torch.float32, device=gt.device)
    call('mmlf_reg_to_class', _p(gt), _p(bins_t), S, float(half_step), B, H * W, _p(out), _st())
    return out


@reg_to_class_op.register_fake
def _(gt, bins_t, half_step):
    B, H, W = gt.shape
    return gt.new_empty((B, bins_t.numel(), H, W))


@torch.library.custom_op('mmlf::mpi_to_weights', mutates_args=())
def mpi_to_weights_op(mpi: torch.Tensor, bins_t: torch.Tensor, half_step: float) -> torch.Tensor:
    _lib.require_device()
    _chk(mpi, torch.float32, 'mpi')
    B, K, five, H, W = mpi.shape
    S = bins_t.numel()
    out = torch.empty((B, S, H, W), dtype=torch.float32, device=mpi.device)
    call('mmlf_mpi_to_weights', _p(mpi), K, _p(bins_t), S, float(half_step), B, H * W, _p(out), _st())
    return out


@mpi_to_weights_op.register_fake
def _(mpi, bins_t, half_step):
    B, K, five, H, W = mpi.shape
    return mpi.new_empty((B, bins_t.numel(), H, W))


@torch.library.custom_op('mmlf::ese_reduce', mutates_args=())
def ese_reduce(means: torch.Tensor, logvars: torch.Tensor, disp: torch.Tensor) -> list[torch.Tensor]:
    """means/logvars (K, B, H, W) -> [mean, logvar (B, H, W), posterior (B, K, H, W)]  (ensamble.py:78-101)."""
    _lib.require_device()
    _chk(means, torch.float32, 'means')
    _chk(logvars, torch.float32, 'logvars')
    K, B, H, W = means.shape
    mean = torch.empty((B, H, W), dtype=torch.float32, device=means.device)
    logvar = torch.empty_like(mean)
    post = torch.empty((B, K, H, W), dtype=torch.float32, device=means.device)
    call('mmlf_ese_reduce', _p(means), _p(logvars), _p(disp), K, B, H * W, _p(mean), _p(logvar), _p(post), _st())
    return [mean, logvar, post]


@ese_reduce.register_fake
def _(means, logvars, disp):
    K, B, H, W = means.shape
    return [means.new_empty((B, H, W)), means.new_empty((B, H, W)), means.new_empty((B, K, H, W))]


# ----------------------------------------------------------------------------------------- optimiser
@torch.library.custom_op('mmlf::adam_step', mutates_args=('p', 'm', 'v'))
def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1: float,
              beta2: float, eps: float, step: int) -> None:
    _lib.require_device()
    for t, nm in ((p, 'p'), (g, 'g'), (m, 'm'), (v, 'v')):
        _chk(t, torch.float32, nm)
    call('mmlf_adam_step', _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
         int(step), _st())


# ----------------------------------------------------------------------------------------- losses (raw, no autograd)
def zero_(t):
    """Clear a contiguous tensor with cudaMemsetAsync on the current stream (no framework fill kernel)."""
    assert t.is_cuda and t.is_contiguous()
    call('mmlf_zero', _p(t), t.numel() * t.element_size(), _st())
    return t


def zeros_f64(n, device):
    """Zeroed float64 accumulator: cudaMemsetAsync through the C-ABI, no framework fill kernel."""
    t = torch.empty(n, dtype=torch.float64, device=device)
    call('mmlf_zero', _p(t), n * 8, _st())
    return t


def _batch_strided(t, name):
    """A (B, H, W) float32 CUDA tensor whose images are dense but may be spaced by any batch stride (a plane of the
    (B, OC, H, W) network output): returned as is with its stride, otherwise made contiguous."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f'mmlf_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)')
    if t.dtype != torch.float32:
        raise RuntimeError(f'mmlf_b200: `{name}` must be torch.float32, got {t.dtype}')
    if t.dim() == 3 and t.stride(2) == 1 and t.stride(1) == t.shape[2] and t.stride(0) >= t.shape[1] * t.shape[2]:
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.numel() // t.shape[0]


def loss_prepass(mask, mask_padding=None, mpi=None):
    """-> double[8] device tensor of normalisers (see include/mmlf_b200.h)."""
    _lib.require_device()
    mask = _chk(mask, torch.int32, 'mask')
    B = mask.shape[0]
    HW = mask.numel() // B
    sums = zeros_f64(8, mask.device)
    K = mpi.shape[1] if mpi is not None else 0
    call('mmlf_loss_prepass', _p(mask), _p(mask_padding), _p(mpi), K, B, HW, _p(sums), _st())
    return sums


def loss_regression(kind, mean, logvar, target, mask, mask_padding, sums, param=0.0, want_grad=True):
    """mean / logvar may be planes of the (B, OC, H, W) network output (``output[:, 0]``): they are read in place through
    their batch stride.  Gradients come back dense (B, H, W)."""
    shape = tuple(mean.shape)
    mean, stride = _batch_strided(mean.reshape(mean.shape[0], -1, mean.shape[-1]) if mean.dim() != 3 else mean, 'mean')
    if logvar is not None:
        logvar, s2 = _batch_strided(logvar.reshape(mean.shape) if logvar.dim() != 3 else logvar, 'logvar')
        if s2 != stride:
            mean, logvar = mean.contiguous(), logvar.contiguous()
            stride = mean.numel() // mean.shape[0]
    target = _chk(target, torch.float32, 'target')
    B = mean.shape[0]
    HW = mean.numel() // B
    K = target.shape[1] if kind in (1, 3) else 0
    loss_sum = zeros_f64(1, mean.device)
    dense = stride == HW
    g_mean = torch.empty(shape, dtype=torch.float32, device=mean.device) if want_grad and kind in (0, 1, 2, 3) else None
    g_logvar = torch.empty(shape, dtype=torch.float32, device=mean.device) if want_grad and kind in (2, 3) else None
    if dense or not want_grad:
        call('mmlf_loss_regression', kind, _p(mean), _p(logvar), _p(target), K, _p(mask), _p(mask_padding), _p(sums),
             float(param), B, HW, _p(loss_sum), _p(g_mean), _p(g_logvar), stride, _st())
    else:
        # strided predictions, dense gradients: the kernel has ONE stride for both, so write the gradients through a
        # tensor with the predictions' stride and hand out its planes
        gbuf = torch.empty((B, stride), dtype=torch.float32, device=mean.device)
        gm = gbuf[:, :HW]
        gl = gbuf[:, HW:2 * HW] if (g_logvar is not None and stride >= 2 * HW) else None
        if g_logvar is not None and gl is None:
            mean, logvar = mean.contiguous(), logvar.contiguous()
            call('mmlf_loss_regression', kind, _p(mean), _p(logvar), _p(target), K, _p(mask), _p(mask_padding), _p(sums),
                 float(param), B, HW, _p(loss_sum), _p(g_mean), _p(g_logvar), HW, _st())
        else:
            call('mmlf_loss_regression', kind, _p(mean), _p(logvar), _p(target), K, _p(mask), _p(mask_padding), _p(sums),
                 float(param), B, HW, _p(loss_sum), _p(gm), _p(gl), stride, _st())
            g_mean = gm.view((B,) + shape[1:])
            g_logvar = gl.view((B,) + shape[1:]) if gl is not None else None
    return loss_sum, g_mean, g_logvar


def loss_finish(loss_sum, sums):
    """value = loss_sum / count with no division when the mask is empty (loss.py:73-77), on the device."""
    out = torch.empty(1, dtype=torch.float32, device=loss_sum.device)
    call('mmlf_loss_finish', _p(loss_sum), _p(sums), _p(out), _st())
    return out[0]


def loss_cross_entropy(scores, target, gt, bins_t, half_step, mask, sums, want_grad=True):
    scores = _chk(scores, torch.float32, 'scores')
    B, S, H, W = scores.shape
    loss_sum = zeros_f64(1, scores.device)
    g = torch.empty_like(scores) if want_grad else None
    call('mmlf_loss_cross_entropy', _p(scores), _p(target), _p(gt), _p(bins_t), float(half_step), S, _p(mask),
         _p(sums), B, H * W, _p(loss_sum), _p(g), _st())
    return loss_sum, g
