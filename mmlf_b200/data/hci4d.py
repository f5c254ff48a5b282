"""Drop-in for the hot-path part of ``mmlf.data.hci4d`` (/root/reference/mmlf/data/hci4d.py): view-index
extraction, ``Shift`` / ``RandomShift`` and the texture mask.  Dataset scanning, PNG/PFM I/O and the CPU augmentation
chain are out of scope."""
import math
import random

import numpy as np
import torch

from .. import ops


def create_mask_margin(shape, margin=0):
    """hci4d.py:15-35"""
    assert margin >= 0
    mask = torch.ones(shape, dtype=torch.bool)
    if margin > 0:
        mask[..., :margin, :] = False
        mask[..., -margin:, :] = False
        mask[..., :margin] = False
        mask[..., -margin:] = False
    return mask


def create_mask_texture(center, wsize, threshold):
    """hci4d.py:38-69 on the GPU: one smem-tiled stencil instead of a (1, 3 * wsize^2, H * W) unfold.  center: CUDA
    tensor (B, 3, H, W); returns the int mask (B, H, W) with a margin of wsize // 2 zeroed."""
    return ops.texture_mask(center, wsize, threshold)


def view_indices(nviews=(9, 9)):
    """hci4d.py:142-149: indices of the horizontal, vertical, rising-diagonal (reversed) and falling-diagonal views
    in the sorted, row-major view list."""
    w, h = nviews
    us = [int(h / 2) * w + i for i in range(h)]
    vs = [int(w / 2) + w * i for i in range(h)]
    ids = [w - i - 1 + w * i for i in range(h)]
    ids.reverse()
    dds = [i + w * i for i in range(h)]
    return us, vs, ids, dds


def extract_stacks(views_u8, n=9):
    """GPU twin of the stack building in ``HCI4D.load_scene`` (hci4d.py:151-193): views_u8 (n*n, H, W, 3) uint8 CUDA
    tensor in sorted-file order -> (h, v, i, d) float32 (n, 3, H, W) + center (3, H, W)."""
    h, v, i, d, center = ops.lf_extract_u8(views_u8.contiguous(), n)
    return h, v, i, d, center


class Shift:
    """hci4d.py:894-990.  CUDA tensors are resampled by one kernel launch (bit-exact); like the reference the
    transform writes its result into the tensors it was given and also subtracts ``disp`` from gt / mpi[:, 4]."""

    def __init__(self, disp):
        assert isinstance(disp, float)
        self.disp = disp

    def __call__(self, data):
        data = list(data)
        h, v, i, d = data[0], data[1], data[2], data[3]          # IndexError for 2 stacks, as hci4d.py:925-926
        if not (isinstance(h, torch.Tensor) and h.is_cuda):
            raise RuntimeError('mmlf_b200.data.hci4d.Shift runs on CUDA tensors only (no CPU fallback); '
                               'CPU-side dataset transforms are out of scope of the B200 hot path')
        oh, ov, oi, od = ops.lf_shift(h.contiguous(), v.contiguous(), i.contiguous(), d.contiguous(), self.disp)
        for dst, src in ((h, oh), (v, ov), (i, oi), (d, od)):
            dst.copy_(src)                                       # in-place semantics of hci4d.py:940-981
        if len(data) > 5:
            data[5] -= float(self.disp)
        if len(data) > 6:
            data[6][:, 4, :, :] -= float(self.disp)
        return tuple(data)


class RandomShift:
    """hci4d.py:993-1028"""

    def __init__(self, disp_range):
        assert isinstance(disp_range, float) or (isinstance(disp_range, tuple) and len(disp_range) == 2)
        self.disp_range = disp_range
        if not isinstance(disp_range, tuple):
            assert disp_range > 0
            self.disp_range = (-disp_range, disp_range)

    def __call__(self, data):
        disp = random.uniform(self.disp_range[0], self.disp_range[1])
        return Shift(disp)(data)
