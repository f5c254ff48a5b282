"""Drop-in for ``mmlf.data.hci4d`` (/root/reference/mmlf/data/hci4d.py): the ``HCI4D`` dataset (scene scanning,
PNG / PFM / MPI loading, ``save_batch``), view-index extraction, ``Shift`` / ``RandomShift`` and the texture mask.

Host work is file I/O only (PNG decode with PIL, PFM / NPZ reads).  Everything that computes runs on the GPU: the 81
uint8 views of a scene are uploaded once and the crosshair stacks come out of ``mmlf_lf_extract_u8`` (bit-exact with the
reference's ``img_as_float(...).astype(float32)``), the texture mask out of ``mmlf_texture_mask``, ``Shift`` out of
``mmlf_lf_shift``.  Scenes are cached in HBM (a 9 x 9 x 512 x 512 scene is 113 MB of stacks), not in host RAM, and the
per-sample augmentation chain of train/cli.py:78-87 is the fused gather of ``mmlf_b200.data.augment.GpuAugmenter`` -- the
per-transform CPU classes (``RandomCrop``, ``RedistColor``, ...) exist upstream only to be composed into that chain and
are not provided one by one."""
import copy
import os
import random

import numpy as np
import torch

from .. import ops
from ..utils import dl, lf, pfm


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


_IMG_EXT = ('.png', '.jpg', '.jpeg')
_NOT_VIEWS = ('normals', 'mask', 'objectids', 'unused', 'edges', 'specular')


def scene_view_files(files):
    """hci4d.py:133-138: the sorted view images of a scene directory listing (auxiliary renders filtered out)."""
    imgs = [f for f in files if f.endswith(_IMG_EXT) and not any(tag in f for tag in _NOT_VIEWS)]
    imgs.sort()
    return imgs


def pick_gt_file(files, center_index):
    """hci4d.py:196-207: the ground-truth PFM of a scene, or None."""
    pfms = [f for f in files if f.endswith('.pfm')]
    if len(pfms) > 1:
        pfms = [f for f in pfms if 'disp' in f]
    if len(pfms) > 1:
        pfms = [f for f in pfms if 'lowres' in f]
    if len(pfms) > 1:
        pfms = [f for f in pfms if str(center_index).zfill(3) in f]
    return pfms[0] if pfms else None


def _read_rgb_u8(fname):
    from PIL import Image
    with Image.open(fname) as im:
        if im.mode not in ('RGB', 'RGBA', 'L', 'P'):
            raise ValueError(f'{fname}: unsupported image mode {im.mode} (8-bit RGB views expected)')
        return np.asarray(im.convert('RGB'))


class HCI4D:
    """hci4d.py:72-413 -- the synthetic HCI 4D Light Field Dataset (one sub-directory per scene holding the
    ``input_Cam*.png`` views, ``gt_disp_lowres.pfm``, optionally ``gt_mpi_lowres.npz`` and ``mask.png``).

    Same constructor and item layout as the reference, (h_views, v_views, i_views, d_views, center, gt, mpi, mask, index);
    the items are CUDA tensors on ``device`` (extension keyword) except ``index`` (numpy, as upstream), so use
    ``DataLoader(..., num_workers=0)`` or iterate directly.  ``cache=True`` keeps the scenes resident in HBM.  Scene
    directories are visited in sorted order (upstream: ``os.scandir`` order)."""

    def __init__(self, root, nviews=(9, 9), transform=None, cache=False, length=0, load_dict=False, device='cuda'):
        if load_dict:
            raise NotImplementedError('load_dict (data_k.mat dictionaries) belongs to the retired INN variant')
        self.name = os.path.basename(root)
        entries = sorted((f for f in os.scandir(root) if f.is_dir()), key=lambda f: f.name)
        self.scenes_names = [f.name for f in entries]
        self.scenes = [f.path for f in entries]
        if not self.scenes:
            raise FileNotFoundError(f'HCI4D: no scene directories under {root!r}')
        self.nviews = nviews
        self.transform = transform
        self.length = length
        self.device = torch.device(device)
        self.cache = cache
        if cache:
            self.data = []
            self.cache_scenes()

    def load_scene(self, index):
        """hci4d.py:124-254."""
        scene = self.scenes[index]
        files = [f.name for f in os.scandir(scene)]
        imgs = scene_view_files(files)
        w, h = self.nviews
        if w != h:
            raise NotImplementedError('the crosshair extraction kernel takes square view grids (nviews = (n, n))')
        if len(imgs) < w * h:
            raise FileNotFoundError(f'{scene}: {len(imgs)} view images, {w * h} expected')
        us, vs, ids, dds = view_indices(self.nviews)
        # host: decode only the views the four stacks read (33 of 81), into a pinned (n * n, H, W, 3) uint8 array
        first = _read_rgb_u8(os.path.join(scene, imgs[us[0]]))
        H, W = first.shape[:2]
        host = torch.zeros((w * h, H, W, 3), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else \
            torch.zeros((w * h, H, W, 3), dtype=torch.uint8)
        for j in sorted(set(us + vs + ids + dds)):
            img = first if j == us[0] else _read_rgb_u8(os.path.join(scene, imgs[j]))
            if img.shape != (H, W, 3):
                raise ValueError(f'{scene}/{imgs[j]}: view size {img.shape} differs from {(H, W, 3)}')
            host[j] = torch.from_numpy(img)
        dev = self.device
        h_views, v_views, i_views, d_views, center = extract_stacks(host.to(dev, non_blocking=True), w)
        gt_file = pick_gt_file(files, us[int(w / 2)])
        if gt_file is not None:
            gt_np = np.flip(pfm.load(os.path.join(scene, gt_file)), 0).copy()          # hci4d.py:210-213
            gt = torch.from_numpy(np.ascontiguousarray(gt_np, dtype=np.float32)).to(dev)
        else:
            gt = torch.zeros((H, W), dtype=torch.float32, device=dev)
        if 'gt_mpi_lowres.npz' in files:                                               # hci4d.py:216-222
            mpi_np = np.load(os.path.join(scene, 'gt_mpi_lowres.npz'))['mpi']
            mpi_np = np.flip(mpi_np, 0).copy().transpose((2, 3, 0, 1))
            mpi_np[np.isnan(mpi_np)] = 0.0
            mpi = torch.from_numpy(np.ascontiguousarray(mpi_np[:12])).to(dev)
        else:                                                                          # one plane: centre view + gt
            mpi = torch.zeros((1, 5, H, W), dtype=torch.float64, device=dev)
            mpi[0, :3], mpi[0, 3], mpi[0, 4] = center, 1.0, gt
        fname = os.path.join(scene, 'mask.png')
        if os.path.exists(fname):
            mask = torch.from_numpy((_read_rgb_u8(fname)[:, :, 0] > 0).astype(np.int64)).to(dev)
        else:
            mask = torch.ones((H, W), dtype=torch.int64, device=dev)
        mask = mask * create_mask_texture(center.unsqueeze(0), 23, 0.02)[0].long()      # hci4d.py:241-243
        return h_views, v_views, i_views, d_views, center, gt, mpi, mask, np.atleast_1d(index)

    def cache_scenes(self):
        print('Caching dataset "{}"...'.format(self.name))
        for i in range(len(self.scenes)):
            self.data.append(self.load_scene(i))

    def __len__(self):
        return len(self.scenes) if self.length == 0 else self.length

    def __getitem__(self, index):
        index = index % len(self.scenes)
        data = self.data[index] if self.cache else self.load_scene(index)
        if self.transform:
            data = tuple(t.clone() if isinstance(t, torch.Tensor) else copy.deepcopy(t) for t in data)
            data = self.transform(data)
        return data

    def save_batch(self, path, index, result=None, uncert=None, runtime=None, gmm=None, nll=None, posterior=None):
        """hci4d.py:295-413: ``scenes/<scene>/`` (views, center, gt, diff, result / uncert as PNG + PFM, gmm / nll /
        posterior as NPY) and ``ours/disp_maps/<scene>.pfm``, ``ours/runtimes/<scene>.txt``.  numpy inputs, as upstream."""
        scenes = os.path.join(path, 'scenes')
        disp_maps = os.path.join(path, 'ours', 'disp_maps')
        runtimes = os.path.join(path, 'ours', 'runtimes')
        for d in (scenes, disp_maps, runtimes):
            os.makedirs(d, exist_ok=True)
        host = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)  # noqa: E731
        for arr_i, i in enumerate(np.asarray(index).squeeze(1).tolist()):
            i = int(i)
            scene = self.scenes_names[i]
            scene_dir = os.path.join(scenes, scene)
            h_views, v_views, i_views, d_views, center, gt, mpi, mask, _ = [
                host(t) for t in self.__getitem__(i)]
            lf.save_views(scene_dir, h_views, v_views, i_views, d_views)
            dl.save_img(os.path.join(scene_dir, 'center.png'), center)
            dl.save_img(os.path.join(scene_dir, 'gt.png'), gt)
            if result is not None:
                dl.save_img(os.path.join(scene_dir, 'diff.png'), np.abs(gt - result[arr_i]))
            pfm.save(os.path.join(scene_dir, 'gt.pfm'), np.flip(gt, 0).astype(np.float32))
            if result is not None:
                res_out = np.flip(np.asarray(result[arr_i], np.float32), 0).copy()
                pfm.save(os.path.join(scene_dir, 'result.pfm'), res_out)
                pfm.save(os.path.join(disp_maps, f'{scene}.pfm'), res_out)
                lo, hi = float(np.min(gt)), float(np.max(gt))
                span = (hi - lo) if hi > lo else 1.0
                dl.save_img(os.path.join(scene_dir, 'result.png'), np.clip((result[arr_i] - lo) / span, 0.0, 1.0))
            if uncert is not None:
                pfm.save(os.path.join(scene_dir, 'uncert.pfm'), np.flip(np.asarray(uncert[arr_i], np.float32), 0).copy())
                dl.save_img(os.path.join(scene_dir, 'uncert.png'), uncert[arr_i])
            if gmm is not None:
                np.save(os.path.join(scene_dir, 'gmm.npy'), gmm[:, :, arr_i])
            if nll is not None:
                np.save(os.path.join(scene_dir, 'nll.npy'), nll[arr_i, ...])
            if posterior is not None:
                np.save(os.path.join(scene_dir, 'posterior.npy'), posterior[arr_i, ...])
            if runtime is not None:
                with open(os.path.join(runtimes, f'{scene}.txt'), 'w') as f:
                    f.write(str(runtime / float(np.asarray(index).shape[0])))


def extract_stacks(views_u8, n=9):
    """GPU twin of the stack building in ``HCI4D.load_scene`` (hci4d.py:151-193): views_u8 (n*n, H, W, 3) uint8 CUDA
    tensor in sorted-file order -> (h, v, i, d) float32 (n, 3, H, W) + center (3, H, W)."""
    h, v, i, d, center = ops.lf_extract_u8(views_u8.contiguous(), n)
    return h, v, i, d, center


class Shift:
    """hci4d.py:894-990.  The four stacks are resampled by one kernel launch (bit-exact); like the reference the
    transform writes its result into the arrays it was given and also subtracts ``disp`` from gt / mpi[:, 4].

    CUDA tensors are processed where they are.  numpy arrays and CPU tensors (the reference uses ``Shift`` as a CPU dataset
    transform, train/cli.py:89-90) are staged through the GPU: upload, the same kernel, download into the caller's
    arrays -- there is no CPU implementation of the resampling in this package."""

    def __init__(self, disp):
        assert isinstance(disp, float)
        self.disp = disp

    def __call__(self, data):
        data = list(data)
        stacks = [data[0], data[1], data[2], data[3]]            # IndexError for 2 stacks, as hci4d.py:925-926
        dev_in = []
        for t in stacks:
            if isinstance(t, np.ndarray):
                if t.dtype != np.float32:
                    raise TypeError('Shift: view stacks must be float32 (hci4d.py:155-156)')
                dev_in.append(torch.from_numpy(np.ascontiguousarray(t)).cuda())
            elif not t.is_cuda:
                dev_in.append(t.contiguous().cuda())
            else:
                dev_in.append(t.contiguous())
        outs = ops.lf_shift(*dev_in, self.disp)
        for dst, src in zip(stacks, outs):                       # in-place semantics of hci4d.py:940-981
            if isinstance(dst, np.ndarray):
                dst[...] = src.cpu().numpy()
            else:
                dst.copy_(src)
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
