"""Synthetic light-field source for the CLIs.  The reference's dataset loader (HCI4D: PNG/PFM scanning, skimage, the CPU
augmentation chain) is out of scope of the B200 hot path and the HCI data is not available
offline, so the drop-in CLIs run on seeded synthetic scenes shaped exactly like ``HCI4D.__getitem__``'s 9-tuple
(h_views, v_views, i_views, d_views, center, gt, mpi, mask, index) -- hci4d.py:250-254."""
import numpy as np
import torch


def _scene(seed, n, H, W):
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing='ij')
    d = 1.5 * np.sin(2 * np.pi * (yy / (1.7 * H) + xx / (2.3 * W))) + 0.8 * (xx > 0.55 * W)
    f = rng.uniform(0.05, 0.9, (3, 4, 2)).astype(np.float32)
    ph = rng.uniform(0, 6.28, (3, 4)).astype(np.float32)
    c = n // 2
    stacks = np.zeros((4, n, 3, H, W), np.float32)
    offs = [[(0, k - c) for k in range(n)], [(k - c, 0) for k in range(n)],
            [(c - k, k - c) for k in range(n)], [(k - c, k - c) for k in range(n)]]
    for s in range(4):
        for k, (dv, du) in enumerate(offs[s]):
            ys, xs = yy + d * dv, xx + d * du
            for ch in range(3):
                t = sum(np.sin(f[ch, j, 0] * ys + f[ch, j, 1] * xs + ph[ch, j]) for j in range(4))
                stacks[s, k, ch] = 0.5 + 0.125 * t
    return stacks, d.astype(np.float32)


class SyntheticLF(torch.utils.data.Dataset):
    """``length`` scenes of ``n`` x ``n`` views, H x W pixels; items are the reference's 9-tuple (numpy arrays)."""

    def __init__(self, length=8, n=9, H=64, W=64, seed=0, name='synthetic'):
        self.length, self.n, self.H, self.W, self.seed, self.name = length, n, H, W, seed, name
        self.scenes = [f'{name}_{i:03d}' for i in range(length)]

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        index = index % self.length
        st, gt = _scene(self.seed * 1000 + index, self.n, self.H, self.W)
        center = st[1][self.n // 2].copy()
        mpi = np.zeros((1, 5, self.H, self.W), np.float32)
        mpi[0, :3], mpi[0, 3], mpi[0, 4] = center, 1.0, gt
        mask = np.ones_like(gt, dtype=np.int64)
        return st[0], st[1], st[2], st[3], center, gt, mpi, mask, np.atleast_1d(index)
