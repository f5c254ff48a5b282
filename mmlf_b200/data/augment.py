"""GPU twin of the training augmentation chain of the reference.

The reference composes, per sample and in CPU dataset workers (/root/reference/mmlf/train/cli.py:78-87),

    RandomDownSampling -> RandomShift -> RandomCrop(ps + 16) -> CenterCrop(ps) -> RandomRotate -> RedistColor ->
    Brightness -> Contrast                       (/root/reference/mmlf/data/hci4d.py:483-530, 894-1028, 533-664,
                                                  1031-1087, 667-785)

and ships 2 GB of float32 patches per 512-patch step to the GPU.  Here the (static-shifted) scenes stay resident in HBM,
the host only draws the random parameters -- with Python's ``random`` in exactly the reference's order, so a seeded run
visits the same samples -- and one gather kernel evaluates the whole chain per output pixel.
"""
import ctypes as C
import random as _random

import numpy as np
import torch

from .. import _lib
from .._lib import call


class AugSample(C.Structure):
    """Mirror of ``mmlf_aug_sample`` (include/mmlf_b200.h)."""
    _fields_ = [('scene', C.c_int), ('f', C.c_int), ('cy', C.c_int), ('cx', C.c_int), ('r', C.c_int),
                ('src', C.c_int * 4), ('flip', C.c_int * 4), ('w0', C.c_float * 16), ('w1', C.c_float * 16),
                ('s0', C.c_int * 16), ('s1', C.c_int * 16), ('mat', C.c_double * 9), ('bright', C.c_float),
                ('contrast', C.c_float), ('one_minus_contrast', C.c_float), ('disp_f', C.c_float), ('disp', C.c_double)]


def draw_params(rng, H, W, ps, max_factor=4, shift_range=1.0, level=0.9):
    """Random parameters of one sample, drawn from ``rng`` (``random`` or a ``random.Random``) in the order in which the
    reference's transforms call it: hci4d.py:526, 1024, 659-660, 1082, 683-690, 774, 739."""
    f = rng.randint(1, max_factor)
    disp = rng.uniform(-shift_range, shift_range)
    hd, wd = -(-H // f), -(-W // f)
    size = ps + 2 * 4 * 2
    assert hd > size and wd > size, 'patch + margin does not fit the down-sampled scene (hci4d.py:656-657)'
    y = rng.randint(0, hd - size)
    x = rng.randint(0, wd - size)
    r = rng.randint(0, 3)
    m = np.zeros((3, 3))
    m[0, 0] = rng.uniform(0.0, 1.0)
    m[0, 1] = rng.uniform(0.0, 1.0 - m[0, 0])
    m[1, 0] = rng.uniform(0.0, 1.0 - m[0, 0])
    m[1, 1] = rng.uniform(0.0, 1.0 - max(m[0, 1], m[1, 0]))
    m[0, 2] = 1.0 - m[0, 0] - m[0, 1]
    m[1, 2] = 1.0 - m[1, 0] - m[1, 1]
    m[2, 0] = 1.0 - m[0, 0] - m[1, 0]
    m[2, 1] = 1.0 - m[0, 1] - m[1, 1]
    m[2, 2] = m[0, 0] + m[0, 1] + m[1, 0] + m[1, 1] - 1.0
    bright = rng.uniform(-level, level) + 1.0
    contrast = rng.uniform(-level, level) + 1.0
    return dict(f=f, disp=disp, y=y, x=x, r=r, mat=m, bright=bright, contrast=contrast, ps=ps)


def draw_params_plain(rng, H, W, ps):
    """``--train_no_data_augment`` (train/cli.py:72-76): RandomCrop(ps + 16) -> CenterCrop(ps) only; every other stage of
    the fused chain gets its identity parameters (factor 1, zero shift, no rotation, identity colour matrix, gains of 1 --
    the Contrast pass is skipped by the caller, it is not an exact identity in float32)."""
    size = ps + 2 * 4 * 2
    assert H > size and W > size, 'patch + margin does not fit the scene (hci4d.py:656-657)'
    y = rng.randint(0, H - size)
    x = rng.randint(0, W - size)
    return dict(f=1, disp=0.0, y=y, x=x, r=0, mat=np.eye(3), bright=1.0, contrast=1.0, ps=ps)


def pack_samples(params, scene_ids, n):
    """list of draw_params() dicts (+ scene index each) -> ctypes array of mmlf_aug_sample (host memory)."""
    arr = (AugSample * len(params))()
    lib = _lib.lib()
    for s, p, sc in zip(arr, params, scene_ids):
        s.scene, s.f = int(sc), int(p['f'])
        s.cy, s.cx = int(p['y']) + 8, int(p['x']) + 8          # RandomCrop origin + CenterCrop margin (16 / 2)
        rc = lib.mmlf_augment_fill(C.byref(s), int(p['r']), float(p['disp']), int(n))
        if rc:
            raise RuntimeError(lib.mmlf_last_error().decode())
        for j in range(9):
            s.mat[j] = float(p['mat'].reshape(-1)[j])
        s.bright = float(p['bright'])
        s.contrast = float(p['contrast'])
        s.one_minus_contrast = float(1.0 - p['contrast'])
    return arr


class GpuAugmenter:
    """Scenes resident on the GPU + the fused augmentation chain.

    scenes: list of load_scene-style 9-tuples (h, v, i, d, center, gt, mpi, mask, index) as numpy arrays or tensors, all of
    one size; a static ``Shift(train_shift)`` (train/cli.py:89-90) must already have been applied to them.
    """

    def __init__(self, scenes, device='cuda'):
        _lib.require_device()
        T = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).to(device=device, dtype=dt)  # noqa: E731
        self.stacks = torch.stack([torch.stack([T(s[k], torch.float32) for k in range(4)]) for s in scenes]).to(device)
        self.center = torch.stack([T(s[4], torch.float32) for s in scenes]).to(device).contiguous()
        self.gt = torch.stack([T(s[5], torch.float32) for s in scenes]).to(device).contiguous()
        self.mpi = torch.stack([T(s[6], torch.float32) for s in scenes]).to(device).contiguous()
        self.mask = torch.stack([T(s[7], torch.int32) for s in scenes]).to(device).contiguous()
        self.index = [np.atleast_1d(s[8]) for s in scenes]
        self.stacks = self.stacks.contiguous()
        self.S, _, self.n, _, self.H, self.W = self.stacks.shape
        self.K = self.mpi.shape[1]
        self.device = self.stacks.device

    def draw(self, B, ps, rng=_random, max_factor=4, plain=False):
        """B scene indices + parameter sets from the host RNG (scene first, as the DataLoader picks the item first)."""
        ids = [rng.randrange(self.S) for _ in range(B)]
        if plain:
            return ids, [draw_params_plain(rng, self.H, self.W, ps) for _ in range(B)]
        return ids, [draw_params(rng, self.H, self.W, ps, max_factor) for _ in range(B)]

    def pack(self, scene_ids, params):
        """Host side of one batch: the parameter records of its samples, uploaded -> device tensor for ``run``."""
        host = pack_samples(params, scene_ids, self.n)
        return torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(self.device)

    def run(self, samples, B, ps, mean_override=None, plain=False):
        """Device side: the two kernels over packed sample records (no host work besides the launches)."""
        n, dev = self.n, self.device
        views = torch.empty((4, B, n, 3, ps, ps), dtype=torch.float32, device=dev)
        center = torch.empty((B, 3, ps, ps), dtype=torch.float32, device=dev)
        sums = torch.zeros(B, dtype=torch.float64, device=dev)
        gt = torch.empty((B, ps, ps), dtype=torch.float32, device=dev)
        mpi = torch.empty((B, self.K, 5, ps, ps), dtype=torch.float32, device=dev)
        mask = torch.empty((B, ps, ps), dtype=torch.int32, device=dev)
        P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        call('mmlf_augment_patches', P(self.stacks), P(self.center), P(self.gt), P(self.mpi), P(self.mask), self.S, n,
             self.K, self.H, self.W, P(samples), B, ps, P(views), P(center), P(sums), P(gt), P(mpi), P(mask), st)
        mo = None
        if mean_override is not None:
            mo = torch.as_tensor(np.asarray(mean_override, dtype=np.float32)).to(dev)
        if not plain:
            call('mmlf_augment_contrast', P(views), P(center), P(samples), P(sums), P(mo), B, n, ps, st)
        self.last_means = sums / float(n * 3 * ps * ps)
        return views, center, gt, mpi, mask

    def __call__(self, scene_ids, params, mean_override=None, plain=False):
        """-> (h, v, i, d (B, n, 3, ps, ps), center (B, 3, ps, ps), gt (B, ps, ps), mpi (B, K, 5, ps, ps) f32,
        mask (B, ps, ps) int32, index (B, 1)) on the GPU."""
        B, ps = len(params), int(params[0]['ps'])
        views, center, gt, mpi, mask = self.run(self.pack(scene_ids, params), B, ps, mean_override, plain)
        index = torch.from_numpy(np.stack([self.index[i] for i in scene_ids]))
        return views[0], views[1], views[2], views[3], center, gt, mpi, mask, index
