"""Drop-in for ``mmlf.model.loss`` (/root/reference/mmlf/model/loss.py): same classes, same call signatures,
``forward(input: dict, target, mask[, mask_padding]) -> 0-d tensor``.

Each loss is one pre-pass kernel (global normalisers, all-reduced across ranks when running data-parallel, SURVEY.md
H4) plus one fused value+gradient kernel; the gradient is handed to autograd by a custom Function.
"""
import torch
import torch.nn as nn

from .. import ops, parallel


def create_mask_margin(shape, margin=0):
    """loss.py:6-26 (host-side helper, identical semantics)."""
    assert margin >= 0
    mask = torch.ones(shape, dtype=torch.bool)
    if margin > 0:
        mask[..., :margin, :] = False
        mask[..., -margin:, :] = False
        mask[..., :margin] = False
        mask[..., -margin:] = False
    return mask


def _as_i32(mask):
    # loss.py:71-72: `mask.int().sum()` / `mask.float()`: any integer / bool mask is accepted
    if mask is None:
        return None
    return mask.contiguous() if mask.dtype == torch.int32 else mask.to(torch.int32).contiguous()


def _as_f32(t):
    # the synthetic 1-plane MPI of the reference is float64; kernels compute in f32
    return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()


def _finish(loss_sum, sums):
    """value = sum / count, with no division when count == 0 (loss.py:73-77) -- without a host sync."""
    return ops.loss_finish(loss_sum, sums)


class _RegressionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, mean, logvar, target, mask, mask_padding, param):
        mask = _as_i32(mask)
        mask_padding = _as_i32(mask_padding)
        target = _as_f32(target)
        sums = ops.loss_prepass(mask, mask_padding, target if kind in (1, 3) else None)
        parallel.all_reduce_sum_(sums)
        want = mean.requires_grad or (logvar is not None and logvar.requires_grad)
        loss_sum, g_mean, g_logvar = ops.loss_regression(kind, mean, logvar, target, mask, mask_padding, sums, param,
                                                         want_grad=want)
        parallel.all_reduce_sum_(loss_sum)
        ctx.save_for_backward(g_mean, g_logvar)
        ctx.shapes = (mean.shape, None if logvar is None else logvar.shape)
        return _finish(loss_sum, sums)

    @staticmethod
    def backward(ctx, g):
        g_mean, g_logvar = ctx.saved_tensors
        gm = None if g_mean is None else (g_mean * g).view(ctx.shapes[0])
        gl = None if g_logvar is None else (g_logvar * g).view(ctx.shapes[1])
        return None, gm, gl, None, None, None, None


class _CrossEntropyLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, target, mask):
        mask = _as_i32(mask)
        target = _as_f32(target)
        sums = ops.loss_prepass(mask)
        parallel.all_reduce_sum_(sums)
        loss_sum, g = ops.loss_cross_entropy(scores.contiguous(), target, None, None, 0.0, mask, sums,
                                             want_grad=scores.requires_grad)
        parallel.all_reduce_sum_(loss_sum)
        ctx.save_for_backward(g)
        return _finish(loss_sum, sums)

    @staticmethod
    def backward(ctx, gscalar):
        (g,) = ctx.saved_tensors
        return (None if g is None else g * gscalar), None, None


def _value_only(kind, input, target, mask, param=0.0):
    mask = _as_i32(mask)
    sums = ops.loss_prepass(mask)
    parallel.all_reduce_sum_(sums)
    mean = input['mean'].detach()
    loss_sum, _, _ = ops.loss_regression(kind, mean, None, _as_f32(target).reshape(mean.shape), mask, None, sums,
                                         param, want_grad=False)
    parallel.all_reduce_sum_(loss_sum)
    return _finish(loss_sum, sums)


class MaskedL1Loss(nn.Module):
    """loss.py:29-77"""

    def forward(self, input, target, mask):
        mean = input['mean']
        return _RegressionLoss.apply(0, mean, None, target.reshape(mean.shape), mask, None, 0.0)


class MultiMaskedL1Loss(nn.Module):
    """loss.py:80-103"""

    def forward(self, input, target, mask):
        return _RegressionLoss.apply(1, input['mean'], None, target, mask, None, 0.0)


class MaskedMSELoss(nn.Module):
    """loss.py:106-122 (validation metric)"""

    def forward(self, input, target, mask):
        return _value_only(4, input, target, mask)


class MaskedBadPix(nn.Module):
    """loss.py:163-187 (validation metric)"""

    def __init__(self, t=0.07):
        super(MaskedBadPix, self).__init__()
        self.t = t

    def forward(self, input, target, mask):
        return _value_only(5, input, target, mask, self.t)


class MaskedCrossEntropy(nn.Module):
    """loss.py:137-160"""

    def forward(self, input, target, mask):
        return _CrossEntropyLoss.apply(input['scores'], target, mask)


class ImprovedUncertaintyL1Loss(nn.Module):
    """loss.py:254-294"""

    def forward(self, input, target, mask, mask_padding=None):
        mean = input['mean']
        return _RegressionLoss.apply(2, mean, input['logvar'], target.reshape(mean.shape), mask, mask_padding, 0.0)


class ImprovedMultiUncertaintyL1Loss(nn.Module):
    """loss.py:336-372 (``mask_padding`` accepted and ignored, as in the reference)"""

    def forward(self, input, target, mask, mask_padding=None):
        return _RegressionLoss.apply(3, input['mean'], input['logvar'], target, mask, None, 0.0)
