"""Oracle (test infrastructure): the training augmentation chain with EXPLICIT per-sample parameters.

Restates, on numpy arrays, the transforms that /root/reference/mmlf/train/cli.py:78-87 composes for every sample:

    RandomDownSampling -> RandomShift -> RandomCrop(ps + 16) -> CenterCrop(ps) -> RandomRotate -> RedistColor ->
    Brightness -> Contrast                      (/root/reference/mmlf/data/hci4d.py:483-530, 894-1028, 533-664, 1031-1087,
                                                 667-785)

The reference draws its parameters from Python's global ``random`` inside each transform; here they are arguments
(``draw_params`` replays the reference's draw order), so the same chain can be evaluated by the GPU kernel.  The arithmetic
follows the reference under NumPy >= 2 promotion rules (NEP 50), which is what the goldens were generated with:

  * a Python float next to a float32 array is "weak": the operation stays float32 (DownSampling ``gt /= f``,
    Shift weights, Brightness, Contrast);
  * ``mat[i, j]`` of a float64 ndarray is an np.float64 *scalar*, which is strong: RedistColor multiplies in float64 and
    rounds to float32 when the product is stored / accumulated into the float32 stack (three roundings per channel);
  * the multi-plane image ``mpi`` built by load_scene is float64 (hci4d.py:223-226) and stays float64.
"""
import numpy as np

from .lf import shift as _shift


def draw_params(rng, H, W, ps, max_factor=4, shift_range=1.0, level=0.9):
    """The parameters of one sample in the order the reference draws them from ``random`` (an object with randint /
    uniform, e.g. ``random.Random(seed)`` or the ``random`` module itself)."""
    f = rng.randint(1, max_factor)                                   # RandomDownSampling, hci4d.py:526
    disp = rng.uniform(-shift_range, shift_range)                    # RandomShift, hci4d.py:1024
    hd, wd = -(-H // f), -(-W // f)                                  # shape of x[::f]
    size = ps + 2 * 4 * 2                                            # train/cli.py:80
    assert hd > size and wd > size
    y = rng.randint(0, hd - size)                                    # RandomCrop, hci4d.py:659-660
    x = rng.randint(0, wd - size)
    r = rng.randint(0, 3)                                            # RandomRotate, hci4d.py:1082
    m = np.zeros((3, 3))                                             # RedistColor, hci4d.py:683-697
    m[0, 0] = rng.uniform(0.0, 1.0)
    m[0, 1] = rng.uniform(0.0, 1.0 - m[0, 0])
    m[1, 0] = rng.uniform(0.0, 1.0 - m[0, 0])
    m[1, 1] = rng.uniform(0.0, 1.0 - max(m[0, 1], m[1, 0]))
    m[0, 2] = 1.0 - m[0, 0] - m[0, 1]
    m[1, 2] = 1.0 - m[1, 0] - m[1, 1]
    m[2, 0] = 1.0 - m[0, 0] - m[1, 0]
    m[2, 1] = 1.0 - m[0, 1] - m[1, 1]
    m[2, 2] = m[0, 0] + m[0, 1] + m[1, 0] + m[1, 1] - 1.0
    bright = rng.uniform(-level, level) + 1.0                        # Brightness, hci4d.py:774
    contrast = rng.uniform(-level, level) + 1.0                      # Contrast, hci4d.py:739
    return dict(f=f, disp=disp, y=y, x=x, r=r, mat=m, bright=bright, contrast=contrast, ps=ps)


def _spatial(a):
    return a.ndim >= 2 and a.shape[-1] > 1 and a.shape[-2] > 1


def down_sample(data, f):
    """DownSampling.__call__, hci4d.py:495-510."""
    data = [a[..., ::f, ::f] if _spatial(a) else a for a in data]
    data[5] = data[5] / np.float32(f)                                # float32 array /= python float
    data[6] = data[6].copy()
    data[6][:, 4] /= float(f)                                        # float64 array
    return data


def crop(data, size, y, x):
    """Crop.__call__, hci4d.py:556-577."""
    return [a[..., y:y + size, x:x + size] if _spatial(a) else a for a in data]


def rotate90(data):
    """Rotate90.__call__, hci4d.py:1039-1070: indices 0..6 only (the mask, index 7, is NOT rotated)."""
    data = list(data)
    for i in range(7):
        data[i] = np.flip(np.swapaxes(data[i], -1, -2), -2).copy()
    data[0], data[1] = data[1], data[0]
    data[1] = np.flip(data[1], -4)
    data[2], data[3] = data[3], data[2]
    data[3] = np.flip(data[3], -4)
    return data


def redist_color(data, mat):
    """RedistColor.__call__, hci4d.py:699-716 (float64 products, float32 stores)."""
    data = list(data)
    for i in range(5):
        src = np.ascontiguousarray(data[i])
        out = np.empty_like(src)
        for c in range(3):
            acc = (mat[c, 0] * src[..., 0, :, :].astype(np.float64)).astype(np.float32)
            acc = (acc.astype(np.float64) + mat[c, 1] * src[..., 1, :, :].astype(np.float64)).astype(np.float32)
            acc = (acc.astype(np.float64) + mat[c, 2] * src[..., 2, :, :].astype(np.float64)).astype(np.float32)
            out[..., c, :, :] = acc
        data[i] = out
    return data


def brightness(data, alpha):
    """Brightness.__call__, hci4d.py:776-785."""
    return [a * np.float32(alpha) if i < 5 else a for i, a in enumerate(data)]


def contrast(data, alpha, mean=None):
    """Contrast.__call__, hci4d.py:739-751.  ``mean`` defaults to numpy's float32 (pairwise) mean of data[0]."""
    if mean is None:
        mean = np.ascontiguousarray(data[0]).mean()
    mean = np.float32(mean)
    off = np.float32(mean * np.float32(1.0 - alpha))
    return [a * np.float32(alpha) + off if i < 5 else a for i, a in enumerate(data)], mean


def augment(sample, p, mean=None):
    """sample: the 9-tuple of load_scene (h, v, i, d, center, gt, mpi, mask, index); p: draw_params().
    Returns the transformed 9-tuple (contiguous arrays) and the float32 mean Contrast used."""
    data = [np.asarray(a) for a in sample]
    data = down_sample(data, p['f'])
    sh = _shift(tuple(np.ascontiguousarray(a) for a in data[:4]), float(p['disp']))     # Shift, hci4d.py:907-981
    data[:4] = list(sh)
    data[5] = data[5] - np.float32(p['disp'])                        # hci4d.py:984-985
    data[6] = data[6].copy()
    data[6][:, 4] -= float(p['disp'])                                # hci4d.py:987-988
    data = crop(data, p['ps'] + 16, p['y'], p['x'])
    data = crop(data, p['ps'], 8, 8)                                 # CenterCrop, hci4d.py:606-613
    for _ in range(p['r']):
        data = rotate90(data)
    data = redist_color(data, p['mat'])
    data = brightness(data, p['bright'])
    data, mean = contrast(data, p['contrast'], mean)
    return [np.ascontiguousarray(a) for a in data], mean
