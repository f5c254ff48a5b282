"""Deterministic synthetic inputs shared by oracle/gen_golden.py and the tests.

Follows the recipe of SURVEY.md H1 / section 8(d): structured light fields
(a smooth texture actually displaced by a disparity map), randomised BatchNorm
statistics and scaled conv weights, so that network outputs are not the
degenerate constant that default init + i.i.d. noise produces.
Everything is numpy / CPU-torch with fixed seeds, so the GPU box regenerates
bit-identical inputs without /root/reference.
"""
import numpy as np

FULL_KW = dict(model_ksize=2, model_in_blocks=3, model_out_blocks=8, model_chs=70, model_views=9,
               model_cross=False, model_uncert=False, model_unet=False, model_discrete=False,
               model_no_batchnorm=False, model_batchnorm_momentum=0.1,
               val_disp_min=-3.5, val_disp_max=3.5)


def model_kwargs(variant='base', cross=False, chs=70, **over):
    kw = dict(FULL_KW)
    kw.update(model_chs=chs, model_cross=cross,
              model_uncert=(variant == 'upr'), model_discrete=(variant == 'dpp'))
    kw.update(over)
    return kw


def synth_lf(seed, H, W, n=9, disp_amp=1.5):
    """Returns (views (n, n, 3, H, W) float32 in [0,1], disparity (H, W) float32).

    View (v, u) is the base texture sampled at (y + d*(v-c), x + d*(u-c)) where
    d is a smooth disparity field with one step edge; analytic texture, so the
    warp is exact (no interpolation).
    """
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
    d = disp_amp * np.sin(2 * np.pi * (yy / (1.7 * H) + xx / (2.3 * W)))
    d = d + 0.8 * (xx > 0.55 * W)
    nwave = 6
    fy = rng.uniform(0.05, 0.9, (3, nwave))
    fx = rng.uniform(0.05, 0.9, (3, nwave))
    ph = rng.uniform(0, 2 * np.pi, (3, nwave))
    amp = rng.uniform(0.3, 1.0, (3, nwave))
    c = n // 2
    views = np.zeros((n, n, 3, H, W), np.float64)
    for v in range(n):
        for u in range(n):
            ys = yy + d * (v - c)
            xs = xx + d * (u - c)
            for ch in range(3):
                t = sum(amp[ch, k] * np.sin(fy[ch, k] * ys + fx[ch, k] * xs + ph[ch, k]) for k in range(nwave))
                views[v, u, ch] = 0.5 + 0.5 * t / amp[ch].sum()
    views += rng.uniform(-0.01, 0.01, views.shape)
    return np.clip(views, 0, 1).astype(np.float32), d.astype(np.float32)


def stacks_from_views(views):
    """(n, n, 3, H, W) grid -> h, v, i, d stacks (n, 3, H, W) with the index
    pattern of hci4d.py:142-149 applied to the row-major flattened grid."""
    n = views.shape[0]
    flat = views.reshape((n * n,) + views.shape[2:])
    us = [(n // 2) * n + i for i in range(n)]
    vs = [n // 2 + n * i for i in range(n)]
    ids = [n - i - 1 + n * i for i in range(n)][::-1]
    dds = [i + n * i for i in range(n)]
    return tuple(np.ascontiguousarray(flat[idx]) for idx in (us, vs, ids, dds))


def synth_batch(seed, B, H, W, n=9):
    """B independent light fields -> four stacks (B, n, 3, H, W) + gt (B, H, W)."""
    hs, vs, is_, ds, gts = [], [], [], [], []
    for b in range(B):
        views, d = synth_lf(seed * 1000 + b, H, W, n)
        h, v, i, dd = stacks_from_views(views)
        hs.append(h), vs.append(v), is_.append(i), ds.append(dd), gts.append(d)
    st = lambda x: np.ascontiguousarray(np.stack(x))  # noqa: E731
    return st(hs), st(vs), st(is_), st(ds), st(gts)


def synth_mpi(seed, gt, K=3):
    """(B, K, 5, H, W) float32 multi-plane target [rgb, alpha, disp]; alpha is
    zero on >= 5 % of the pixels (ImprovedMultiUncertaintyL1Loss is NaN otherwise,
    SURVEY.md a19)."""
    rng = np.random.RandomState(seed)
    B, H, W = gt.shape
    mpi = np.zeros((B, K, 5, H, W), np.float32)
    mpi[:, :, :3] = rng.uniform(0, 1, (B, K, 3, H, W))
    alpha = rng.uniform(0, 1, (B, K, H, W)).astype(np.float32)
    alpha[:, 0] = 1.0 - 0.5 * alpha[:, 1]
    dead = rng.uniform(0, 1, (B, 1, H, W)) < 0.08
    alpha = np.where(dead, 0.0, alpha)
    mpi[:, :, 3] = alpha
    off = rng.uniform(-1.5, 1.5, (B, K, 1, 1)).astype(np.float32)
    off[:, 0] = 0
    mpi[:, :, 4] = gt[:, None] + off[..., 0, 0][:, :, None, None]
    return mpi


def synth_mask(seed, B, H, W, margin=3):
    rng = np.random.RandomState(seed)
    m = (rng.uniform(0, 1, (B, H, W)) > 0.1).astype(np.int32)
    if margin:
        m[:, :margin] = 0
        m[:, -margin:] = 0
        m[:, :, :margin] = 0
        m[:, :, -margin:] = 0
    return m


def perturb_state(state, seed, wscale=2.0):
    """In-place on a dict of numpy arrays or torch tensors: scale conv weights,
    randomise BN affine + running statistics.  Works on both
    because it only uses ``*=`` / ``[...] =`` with numpy-generated values."""
    rng = np.random.RandomState(seed)
    for k in sorted(state.keys()):
        v = state[k]
        is_bn = k.split('.')[2] == '3'          # <net>.<block>.<index>.<name>: index 3 is the BatchNorm
        if k.endswith('num_batches_tracked'):
            continue
        shape = tuple(v.shape)

        def put(arr):
            arr = arr.astype(np.float32)
            if isinstance(v, np.ndarray):
                v[...] = arr
            else:
                import torch
                v.copy_(torch.from_numpy(arr))

        if is_bn and k.endswith('running_mean'):
            put(rng.uniform(-0.2, 0.2, shape))
        elif is_bn and k.endswith('running_var'):
            put(rng.uniform(0.5, 1.5, shape))
        elif is_bn and k.endswith('weight'):
            put(rng.uniform(0.6, 1.4, shape))
        elif is_bn and k.endswith('bias'):
            put(rng.uniform(-0.2, 0.2, shape))
        elif k.endswith('weight'):
            cur = v if isinstance(v, np.ndarray) else v.detach().numpy()
            put(cur * wscale)
        else:  # conv bias
            put(rng.uniform(-0.1, 0.1, shape))
    return state


def synth_state(shapes, seed):
    """Deterministic parameters / buffers for a list of (name, shape): He-scaled conv weights, small biases, randomised
    BatchNorm affine and running statistics.  Used for fixtures whose state is too large to store (the U-Net out-net:
    31 M parameters): the generator loads these values into the reference model, the test regenerates them."""
    rng = np.random.RandomState(seed)
    state = {}
    for name, shape in shapes:
        shape = tuple(int(x) for x in shape)
        leaf = name.rsplit('.', 1)[1]
        if leaf == 'num_batches_tracked':
            state[name] = np.zeros(shape, np.int64)
        elif leaf == 'running_mean':
            state[name] = rng.uniform(-0.2, 0.2, shape).astype(np.float32)
        elif leaf == 'running_var':
            state[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif len(shape) == 4:                                        # conv / transposed-conv weight
            fan_in = shape[1] * shape[2] * shape[3] if '.up.' not in name else shape[0]
            state[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif leaf == 'weight':                                       # BatchNorm gamma
            state[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        else:                                                        # biases, BatchNorm beta
            state[name] = rng.uniform(-0.1, 0.1, shape).astype(np.float32)
    return state


# ----------------------------------------------------------------------------- trained-like full-width fixtures
def quantise_delta(w, w_init):
    """int8 quantisation of a trained weight's difference from its seeded initial value: (q int8, scale float32)."""
    delta = (np.asarray(w, np.float32) - np.asarray(w_init, np.float32)).astype(np.float32)
    scale = np.float32(max(float(np.abs(delta).max()), 1e-12) / 127.0)
    q = np.clip(np.rint(delta / scale), -127, 127).astype(np.int8)
    return q, np.array(scale, np.float32)


def trained_trunk_state(trunk, init):
    """Rebuild the trunk state (in-nets + out-net blocks 0..6) of tests/golden/net_trained_trunk.npz:
    conv weights = seeded init + int8 delta * scale (one float32 multiply, one float32 add -> reproducible anywhere),
    everything else stored as is.  ``init``: name -> numpy array of the torch.manual_seed(0) default initialisation."""
    keys = trunk.files if hasattr(trunk, 'files') else trunk.keys()
    state = {}
    for k in keys:
        tag, name = k[:2], k[2:]
        if tag == 'q/':
            d = (trunk[k].astype(np.float32) * np.float32(trunk['s/' + name])).astype(np.float32)
            state[name] = (np.asarray(init[name], np.float32) + d).astype(np.float32)
        elif tag == 'f/':
            state[name] = np.array(trunk[k])
    return state


TRAINED = dict(B=8, ps=32, lr=1e-3, traj_lr=2e-4, traj_steps=20, n_batches=4, seed0=200, grad_stride=13)


def trained_batch(k):
    """Batch k of the trained-like fixtures (same recipe as oracle/gen_golden.py::_trained_batch)."""
    c = TRAINED
    h, v, i, d, gt = synth_batch(c['seed0'] + k, c['B'], c['ps'], c['ps'])
    mask = synth_mask(c['seed0'] + 50 + k, c['B'], c['ps'], c['ps'], margin=3)
    return (h, v, i, d), gt, mask


def trained_state(variant, init, trunk, g):
    """Full state of the trained-like fixture `net_trained_<variant>.npz` (g): trunk + the variant's head and BatchNorm
    statistics.  ``init``: the variant's own torch.manual_seed(0) default init (numpy)."""
    state = {k: np.array(v) for k, v in init.items()}
    state.update(trained_trunk_state(trunk, init))
    for k in g.files:
        if k.startswith('state/'):
            state[k[6:]] = np.array(g[k])
    return state


# ----------------------------------------------------------------------------- on-disk HCI4D-style scenes
def write_pfm(fname, img):
    """Minimal little-endian PFM writer, independent of mmlf_b200.utils.pfm (rows as given: callers store bottom-up)."""
    img = np.ascontiguousarray(img, dtype='<f4')
    with open(fname, 'wb') as f:
        f.write(b'Pf\n%d %d\n-1.000000\n' % (img.shape[1], img.shape[0]))
        img.tofile(f)


def write_hci_scene(scene_dir, seed, H=64, W=64, n=9, with_mask=False, with_mpi=False):
    """One scene in the layout of the HCI 4D Light Field Dataset (hci4d.py:82-95): input_Cam000..080.png (row-major view
    grid), gt_disp_lowres.pfm (stored bottom-up), plus distractor files the loader must skip.  Returns
    (views uint8 (n*n, H, W, 3), gt float32 (H, W))."""
    import os
    from PIL import Image
    os.makedirs(scene_dir, exist_ok=True)
    views, gt = synth_lf(seed, H, W, n)
    u8 = np.clip(np.rint(views * 255.0), 0, 255).astype(np.uint8)              # (n, n, 3, H, W)
    u8 = np.ascontiguousarray(u8.reshape(n * n, 3, H, W).transpose(0, 2, 3, 1))
    for j in range(n * n):
        Image.fromarray(u8[j]).save(os.path.join(scene_dir, f'input_Cam{j:03d}.png'))
    Image.fromarray(u8[0]).save(os.path.join(scene_dir, 'center_normals.png'))    # filtered out by name (hci4d.py:134-137)
    Image.fromarray(u8[1]).save(os.path.join(scene_dir, 'objectids_highres.png'))
    write_pfm(os.path.join(scene_dir, 'gt_disp_lowres.pfm'), gt[::-1])
    write_pfm(os.path.join(scene_dir, 'gt_depth_lowres.pfm'), 2 * gt[::-1])      # loses against the 'disp' file (:199-201)
    if with_mask:
        m = np.zeros((H, W, 3), np.uint8)
        m[4:, :, :] = 255
        Image.fromarray(m).save(os.path.join(scene_dir, 'mask.png'))
    if with_mpi:
        rng = np.random.RandomState(seed + 7)
        mpi = np.zeros((H, W, 3, 5), np.float32)                                 # stored (H, W, K, 5), bottom-up
        mpi[..., :3] = rng.uniform(0, 1, (H, W, 3, 3))
        mpi[..., 0, 3], mpi[..., 1, 3], mpi[..., 2, 3] = 0.7, 0.3, 0.0
        mpi[..., 0, 4] = gt
        mpi[..., 1, 4] = gt + 1.0
        mpi[3, 5, 2, 4] = np.nan                                                 # NaNs become 0 (hci4d.py:220)
        np.savez(os.path.join(scene_dir, 'gt_mpi_lowres.npz'), mpi=mpi[::-1])
    return u8, gt
