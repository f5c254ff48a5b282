"""Portable Float Map I/O -- the file format of the HCI 4D light-field ground truth (`gt_disp_lowres.pfm`) and of the
disparity maps ``HCI4D.save_batch`` writes.  Mirrors the API of ``mmlf.utils.pfm`` (/root/reference/mmlf/utils/pfm.py:6-90:
``load(filename) -> ndarray``, ``save(filename, image, scale=1.0)``); host-side file I/O, no compute.

Format: line 1 ``PF`` (3 channels) or ``Pf`` (1 channel); line 2 ``<width> <height>``; line 3 a scale whose SIGN gives the
byte order (negative = little endian); then height * width (* 3) float32 values, bottom row first (callers flip)."""
import sys

import numpy as np


def _header_line(f):
    line = f.readline()
    if not line:
        raise Exception('Malformed PFM header.')
    return line.decode('ascii', 'replace').strip()


def load(filename):
    """-> float32 array (H, W) or (H, W, 3), rows in file order (utils/pfm.py:6-52)."""
    with open(filename, 'rb') as f:
        magic = _header_line(f)
        if magic not in ('PF', 'Pf'):
            raise Exception('Not a PFM file.')
        dims = _header_line(f).split()
        if len(dims) != 2 or not all(d.isdigit() for d in dims):
            raise Exception('Malformed PFM header.')
        width, height = int(dims[0]), int(dims[1])
        scale = float(_header_line(f))
        data = np.fromfile(f, ('<' if scale < 0 else '>') + 'f4')
    shape = (height, width, 3) if magic == 'PF' else (height, width)
    return np.reshape(data, shape)


def save(filename, image, scale=1.0):
    """float32 (H, W), (H, W, 1) or (H, W, 3) -> file (utils/pfm.py:55-90)."""
    image = np.asarray(image)
    if image.dtype.name != 'float32':
        raise Exception('Image dtype must be float32.')
    if image.ndim == 3 and image.shape[2] == 3:
        magic = b'PF\n'
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b'Pf\n'
    else:
        raise Exception('Image must have H x W x 3, H x W x 1 or H x W dimensions.')
    order = image.dtype.byteorder
    little = order == '<' or (order in '=|' and sys.byteorder == 'little')
    with open(filename, 'wb') as f:
        f.write(magic)
        f.write(b'%d %d\n' % (image.shape[1], image.shape[0]))
        f.write(b'%f\n' % (-scale if little else scale))
        image.tofile(f)
