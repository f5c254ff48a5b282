"""Oracle (test infrastructure): masked losses, metrics, DPP target builders.

Restates /root/reference/mmlf/model/loss.py and the bin helpers of
/root/reference/mmlf/utils/dl.py.  Each loss returns ``(value, grads)`` where
``grads`` maps the network outputs it reads ('mean', 'logvar', 'scores') to
d loss / d output, derived by hand (the reference relies on autograd).
Reductions are carried in float64 and rounded once; the reference reduces in
float32 with PyTorch's pairwise order, so values agree to fp32 round-off.
"""
import numpy as np

from .net import torch_linspace_f32


def create_mask_margin(shape, margin=0):
    """loss.py:6-26 (duplicate at hci4d.py:15-35)."""
    assert margin >= 0
    mask = np.ones(shape, dtype=bool)
    if margin > 0:
        mask[..., :margin, :] = False
        mask[..., -margin:, :] = False
        mask[..., :margin] = False
        mask[..., -margin:] = False
    return mask


def reg_to_class(arr, start, stop, n_steps):
    """utils/dl.py:109-131. arr (B,H,W) -> (B,n_steps,H,W) float32 one-hot
    (window (stop-start)/n_steps wide around torch.linspace bins)."""
    step = (stop - start) / n_steps
    bins = torch_linspace_f32(start, stop, n_steps).reshape(1, -1, 1, 1)
    return (np.abs(bins - arr[:, None]) < step / 2.0).astype(np.float32)


def mpi_to_weights(arr, start, stop, n_steps):
    """utils/dl.py:134-157. arr (B,K,5,H,W): weights=arr[:,:,3], disp=arr[:,:,4]."""
    step = (stop - start) / n_steps
    bins = torch_linspace_f32(start, stop, n_steps).reshape(1, -1, 1, 1, 1)
    weights = arr[:, :, 3][:, None]
    d = arr[:, :, 4][:, None]
    res = (np.abs(bins - d) < step / 2.0).astype(np.float32) * weights
    return res.sum(2)


def class_to_reg(arr, start, stop, n_steps):
    """utils/dl.py:160-182."""
    bins = torch_linspace_f32(start, stop, n_steps).reshape(1, -1, 1, 1)
    return (bins * arr).sum(1, dtype=np.float32)


def _finish(per_px, mask, grads_per_px):
    """Shared tail of every loss: ``loss *= mask.float(); count = mask.sum();
    return loss.sum() / count`` with no division when count == 0 (loss.py:70-77)."""
    m = mask.astype(np.float64)
    count = float(mask.astype(np.int64).sum())
    scale = 1.0 if count == 0 else 1.0 / count
    val = (per_px.astype(np.float64) * m).sum() * scale
    grads = {k: (g.astype(np.float64) * (m if g.ndim == m.ndim else m[:, None]) * scale).astype(np.float32)
             for k, g in grads_per_px.items()}
    return val, grads


def masked_l1(out, target, mask):
    """MaskedL1Loss.forward, loss.py:46-77."""
    d = out['mean'] - target
    return _finish(np.abs(d), mask, {'mean': np.sign(d)})


def masked_mse(out, target, mask):
    """MaskedMSELoss.forward, loss.py:114-122 (validation metric; value only)."""
    d = out['mean'] - target
    return _finish(d * d, mask, {})[0]


def masked_badpix(out, target, mask, t=0.07):
    """MaskedBadPix.forward, loss.py:177-187."""
    bad = (np.abs(out['mean'] - target) > t).astype(np.int64) * mask.astype(np.int64)
    count = int(mask.astype(np.int64).sum())
    return float(bad.sum()) if count == 0 else float(bad.sum()) / count


def multi_masked_l1(out, target, mask):
    """MultiMaskedL1Loss.forward, loss.py:88-103. target (B,K,5,H,W)."""
    w, t = target[:, :, 3], target[:, :, 4]
    d = out['mean'][:, None] - t
    return _finish((np.abs(d) * w).sum(1), mask, {'mean': (np.sign(d) * w).sum(1)})


def masked_cross_entropy(out, target, mask):
    """MaskedCrossEntropy.forward, loss.py:145-160:
    s = relu(scores); l = -log(exp(sum_c s_c t_c) / sum_c exp(s_c))."""
    raw = out['scores'].astype(np.float64)
    s = np.maximum(raw, 0)
    dot = (s * target).sum(1)
    e = np.exp(s)
    z = e.sum(1)
    per_px = np.log(z) - dot
    g = (e / z[:, None] - target) * (raw > 0)
    return _finish(per_px, mask, {'scores': g})


def improved_uncertainty_l1(out, target, mask, mask_padding=None):
    """ImprovedUncertaintyL1Loss.forward, loss.py:262-294."""
    mean, lv = out['mean'].astype(np.float64), out['logvar'].astype(np.float64)
    d = mean - target
    e = np.exp(-lv)
    loss = e * np.abs(d) + lv
    g_mean = e * np.sign(d)
    g_lv = -e * np.abs(d) + 1.0
    if mask_padding is not None:
        mp = mask_padding.astype(np.float64)
        n = float(mp.size)
        k_in = n / mp.sum() if mp.sum() > 0 else 1.0
        mo = 1.0 - mp
        k_oor = n / mo.sum() if mo.sum() > 0 else 1.0
        loss = (loss * mp * k_in + (-lv) * mo * k_oor) / 2.0
        g_mean = g_mean * mp * k_in / 2.0
        g_lv = (g_lv * mp * k_in - mo * k_oor) / 2.0
    return _finish(loss, mask, {'mean': g_mean, 'logvar': g_lv})


def improved_multi_uncertainty_l1(out, target, mask, mask_padding=None):
    """ImprovedMultiUncertaintyL1Loss.forward, loss.py:344-372 (mask_padding is
    accepted and ignored, as in the reference).  NaN when no pixel has
    sum_k w_k < 0.01 (N / 0 * 0), as in the reference."""
    mean, lv = out['mean'].astype(np.float64), out['logvar'].astype(np.float64)
    w, t = target[:, :, 3].astype(np.float64), target[:, :, 4].astype(np.float64)
    d = mean[:, None] - t
    e = np.exp(-lv)
    wsum = w.sum(1)
    mw = wsum.mean()
    loss = ((e[:, None] * np.abs(d) + lv[:, None]) * w).sum(1) / mw
    g_mean = (e[:, None] * np.sign(d) * w).sum(1) / mw
    g_lv = ((-e[:, None] * np.abs(d) + 1.0) * w).sum(1) / mw
    oor = (target[:, :, 3].sum(1) < 0.01).astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        k = np.float64(oor.size) / oor.sum()
        loss = (loss + (-lv) * oor * k) / 2.0
        g_mean = g_mean / 2.0
        g_lv = (g_lv - oor * k) / 2.0
    return _finish(loss, mask, {'mean': g_mean, 'logvar': g_lv})
