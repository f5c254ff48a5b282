"""Oracle (test infrastructure): light-field view extraction and disparity Shift.

Restates, with explicit index arithmetic instead of slice/concatenate:
  * view-index extraction        /root/reference/mmlf/data/hci4d.py:142-193
  * Shift.__call__               /root/reference/mmlf/data/hci4d.py:907-990
  * the ESE shift sweep values   /root/reference/mmlf/model/ensamble.py:61-62
  * create_mask_texture          /root/reference/mmlf/data/hci4d.py:38-69
"""
import math

import numpy as np


def view_indices(w=9, h=9):
    """Indices into the lexicographically sorted, row-major w x h view list.

    hci4d.py:142-149: ``us`` centre row, ``vs`` centre column, ``ids`` rising
    diagonal (reversed), ``dds`` falling diagonal.
    """
    us = [int(h / 2) * w + i for i in range(h)]
    vs = [int(w / 2) + w * i for i in range(h)]
    ids = [w - i - 1 + w * i for i in range(h)]
    ids.reverse()
    dds = [i + w * i for i in range(h)]
    return us, vs, ids, dds


def u8_to_f32(img_u8):
    """hci4d.py:156-157: ``skimage.img_as_float(u8)`` (float64 ``x * (1/255)``)
    followed by ``.astype(np.float32)``.  For all 256 inputs this equals the
    correctly rounded f32 quotient x/255 (checked in tests/test_oracle_golden.py).
    """
    return (img_u8.astype(np.float64) * (1.0 / 255.0)).astype(np.float32)


def extract_stacks(views_u8, w=9, h=9):
    """views_u8: (w*h, H, W, 3) uint8 in sorted-file order.

    Returns h_views, v_views, i_views, d_views as (n, 3, H, W) float32 and the
    centre view (3, H, W) = v_views[h // 2]  (hci4d.py:151-193).
    """
    us, vs, ids, dds = view_indices(w, h)
    out = []
    for idx in (us, vs, ids, dds):
        st = np.stack([u8_to_f32(views_u8[i]) for i in idx])
        out.append(np.ascontiguousarray(st.transpose((0, 3, 1, 2))))
    center = out[1][int(h / 2)].copy()
    return out[0], out[1], out[2], out[3], center


def shift_taps(disp, o):
    """Two-tap parameters for the view at offset ``o`` from the centre view.

    hci4d.py:934-938 (and :958-962): ``alpha, shift0 = math.modf(disp * o)``;
    ``alpha = |alpha|``; ``shift1 = shift0 + copysign(1, shift0)``.
    Returns (w0, w1, s0, s1) with the weights already rounded to float32, which
    is how both numpy (weak python scalar) and torch (f32 op-math) apply them.
    """
    alpha, s0 = math.modf(float(disp) * o)
    alpha = abs(alpha)
    s1 = s0 + math.copysign(1.0, s0)
    return np.float32(1.0 - alpha), np.float32(alpha), int(s0), int(s1)


def _src_index(n, s, sign):
    """Source index table of ``cat([x[-s:], x[:-s]])`` (sign=+1, hci4d.py:941-943)
    or ``cat([x[s:], x[:s]])`` (sign=-1, hci4d.py:971-973) along an axis of
    length n.  Python slice clamping makes |s| >= n (and s == 0) the identity.
    """
    j = np.arange(n)
    if s == 0 or abs(s) >= n:
        return j
    return (j - sign * s) % n


def _lerp_axis(x, axis, w0, w1, s0, s1, sign):
    n = x.shape[axis]
    a = np.take(x, _src_index(n, s0, sign), axis=axis)
    b = np.take(x, _src_index(n, s1, sign), axis=axis)
    # two f32 products, one f32 add, no FMA  (hci4d.py:940-945)
    return (a * w0).astype(np.float32) + (b * w1).astype(np.float32)


def shift(data, disp):
    """Out-of-place restatement of ``Shift(disp)(data)`` (hci4d.py:907-990).

    data: sequence (h_views, v_views, i_views, d_views[, center, gt, mpi, ...])
    with the four stacks shaped (..., n, 3, H, W) float32.  Returns a tuple in
    the same order with shifted copies; ``gt`` (index 5) and ``mpi[:, 4]``
    (index 6) get ``disp`` subtracted (hci4d.py:984-988).
    """
    data = [np.array(d, copy=True) if isinstance(d, np.ndarray) else d for d in data]
    hv, vv, iv, dv = data[0], data[1], data[2], data[3]
    w = hv.shape[-4]
    h = vv.shape[-4]
    hw, hh = int(w / 2), int(h / 2)
    for i in range(w):
        w0, w1, s0, s1 = shift_taps(disp, i - hw)
        # along W (last axis): h, i, d stacks  (hci4d.py:940-956)
        for st in (hv, iv, dv):
            st[..., i, :, :, :] = _lerp_axis(st[..., i, :, :, :], -1, w0, w1, s0, s1, +1)
    for i in range(h):
        w0, w1, s0, s1 = shift_taps(disp, i - hh)
        # along H: v and d with the same sign, i with the opposite sign (hci4d.py:964-981)
        vv[..., i, :, :, :] = _lerp_axis(vv[..., i, :, :, :], -2, w0, w1, s0, s1, +1)
        iv[..., i, :, :, :] = _lerp_axis(iv[..., i, :, :, :], -2, w0, w1, s0, s1, -1)
        dv[..., i, :, :, :] = _lerp_axis(dv[..., i, :, :, :], -2, w0, w1, s0, s1, +1)
    if len(data) > 5 and data[5] is not None:
        data[5] = data[5] - float(disp)
    if len(data) > 6 and data[6] is not None:
        m = np.array(data[6], copy=True)
        m[:, 4, :, :] -= float(disp)
        data[6] = m
    return tuple(data)


def ese_shift_values(disp_min, disp_max, disp_step):
    """ensamble.py:61-62: the members are ``np.arange`` values in float64, with
    its round-off (member 35 of the default sweep is 3.1e-15, not 0)."""
    return [float(v) for v in np.arange(disp_min, disp_max, disp_step)]


def texture_mae(center, wsize):
    """hci4d.py:53-58: mean over the 3 colours and the wsize x wsize window (zero padded: ``unfold(padding=wsize//2)``)
    of |neighbour - centre pixel|.  center (B, 3, H, W) float32 -> (B, H, W) float32 (accumulated in float64)."""
    B, C, H, W = center.shape
    r = wsize // 2
    pad = np.zeros((B, C, H + 2 * r, W + 2 * r), np.float64)
    pad[:, :, r:r + H, r:r + W] = center
    acc = np.zeros((B, H, W), np.float64)
    c64 = center.astype(np.float64)
    for dy in range(wsize):
        for dx in range(wsize):
            acc += np.abs(pad[:, :, dy:dy + H, dx:dx + W] - c64).sum(1)
    return (acc / (C * wsize * wsize)).astype(np.float32)


def create_mask_texture(center, wsize, threshold):
    """hci4d.py:38-69: (mean L1 >= threshold) with a margin of wsize // 2 masked out; int32 (B, H, W)."""
    mask = (texture_mae(center, wsize) >= np.float32(threshold)).astype(np.int32)
    r = wsize // 2
    if r > 0:
        mask[..., :r, :] = 0
        mask[..., -r:, :] = 0
        mask[..., :r] = 0
        mask[..., -r:] = 0
    return mask
