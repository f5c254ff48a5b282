"""Oracle (test infrastructure): the U-Net out-net of ``--model_unet``.

Restates /root/reference/mmlf/model/unet.py:8-132 as used by feed_forward.py:189-204 (``UNet(chs, out_chs, depth=5,
padding=True, batch_norm=True)``, up_mode 'upconv') on NHWC numpy arrays, forward and hand-derived backward:

  down path i = 0..4 : [conv3x3 p1 -> ReLU -> BN -> conv3x3 p1 -> ReLU -> BN] (64 * 2^i channels), 2x2 max-pool between
  up path            : ConvTranspose2d(k 2, stride 2) -> cat([up, center_crop(bridge)]) -> the same conv block
  last               : 1x1 conv to 1 (BASE) / 2 (UPR) channels

Note the block order conv -> ReLU -> BN (unet.py:85-96), unlike the plain out-net's conv -> BN -> ReLU, and that these
BatchNorm layers use PyTorch's default momentum 0.1 (unet.py:89), not --model_batchnorm_momentum.
SURVEY.md section 8f.4 groundwork: no CUDA path uses this yet.
"""
import numpy as np

from .net import conv2x2 as conv, conv2x2_bwd as conv_bwd

EPS = 1e-5
MOMENTUM = 0.1


def _bn_fwd(p, name, a, training):
    g, b = p[name + '.weight'], p[name + '.bias']
    if training:
        a2 = a.reshape(-1, a.shape[-1]).astype(np.float64)
        n = a2.shape[0]
        mean, var = a2.mean(0), a2.var(0)
        p[name + '.running_mean'] = ((1 - MOMENTUM) * p[name + '.running_mean'] + MOMENTUM * mean).astype(np.float32)
        p[name + '.running_var'] = ((1 - MOMENTUM) * p[name + '.running_var'] + MOMENTUM * var * n / max(n - 1, 1)).astype(np.float32)
        p[name + '.num_batches_tracked'] = p[name + '.num_batches_tracked'] + 1
        mean, var = mean.astype(np.float32), var.astype(np.float32)
    else:
        mean, var = p[name + '.running_mean'], p[name + '.running_var']
    invstd = (1.0 / np.sqrt(var.astype(np.float64) + EPS)).astype(np.float32)
    xhat = (a - mean) * invstd
    return xhat * g + b, {'xhat': xhat, 'invstd': invstd, 'name': name}


def _bn_bwd(p, rec, gy, training):
    g = p[rec['name'] + '.weight']
    gy2, xh2 = gy.reshape(-1, gy.shape[-1]), rec['xhat'].reshape(-1, gy.shape[-1])
    grads = {rec['name'] + '.weight': (gy2 * xh2).sum(0), rec['name'] + '.bias': gy2.sum(0)}
    if training:
        n = gy2.shape[0]
        gx = (g * rec['invstd']) * (gy - gy2.mean(0) - rec['xhat'] * (gy2 * xh2).sum(0) / n)
    else:
        gx = gy * (g * rec['invstd'])
    return gx.astype(np.float32), grads


def _block_fwd(p, prefix, x, training):
    """UNetConvBlock (unet.py:80-101): indices 0 conv, 1 ReLU, 2 BN, 3 conv, 4 ReLU, 5 BN."""
    z1 = conv(x, p[prefix + '.0.weight'], p[prefix + '.0.bias'], 1)
    y1, bn1 = _bn_fwd(p, prefix + '.2', np.maximum(z1, 0), training)
    z2 = conv(y1, p[prefix + '.3.weight'], p[prefix + '.3.bias'], 1)
    y2, bn2 = _bn_fwd(p, prefix + '.5', np.maximum(z2, 0), training)
    return y2, {'prefix': prefix, 'x': x, 'z1': z1, 'bn1': bn1, 'y1': y1, 'z2': z2, 'bn2': bn2}


def _block_bwd(p, rec, gy, training):
    prefix, grads = rec['prefix'], {}
    ga2, d = _bn_bwd(p, rec['bn2'], gy, training)
    grads.update(d)
    gy1, gw, gb = conv_bwd(rec['y1'], p[prefix + '.3.weight'], ga2 * (rec['z2'] > 0), 1)
    grads[prefix + '.3.weight'], grads[prefix + '.3.bias'] = gw, gb
    ga1, d = _bn_bwd(p, rec['bn1'], gy1, training)
    grads.update(d)
    gx, gw, gb = conv_bwd(rec['x'], p[prefix + '.0.weight'], ga1 * (rec['z1'] > 0), 1)
    grads[prefix + '.0.weight'], grads[prefix + '.0.bias'] = gw, gb
    return gx, grads


def _pool_fwd(x):
    """F.max_pool2d(x, 2) (unet.py:71): floor division of odd sizes, first maximum wins in the backward pass."""
    B, H, W, C = x.shape
    h, w = H // 2, W // 2
    win = x[:, :2 * h, :2 * w].reshape(B, h, 2, w, 2, C).transpose(0, 1, 3, 5, 2, 4).reshape(B, h, w, C, 4)
    idx = win.argmax(-1)                       # first occurrence, window scanned row-major like PyTorch
    return np.take_along_axis(win, idx[..., None], -1)[..., 0], {'idx': idx, 'shape': x.shape}


def _pool_bwd(rec, gy):
    B, H, W, C = rec['shape']
    h, w = H // 2, W // 2
    gwin = np.zeros((B, h, w, C, 4), np.float32)
    np.put_along_axis(gwin, rec['idx'][..., None], gy[..., None], -1)
    gx = np.zeros(rec['shape'], np.float32)
    gx[:, :2 * h, :2 * w] = gwin.reshape(B, h, w, C, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(B, 2 * h, 2 * w, C)
    return gx


def _upconv_fwd(x, w, b):
    """nn.ConvTranspose2d(cin, cout, 2, stride=2) (unet.py:107-108); w: (cin, cout, 2, 2)."""
    B, H, W, C = x.shape
    out = np.zeros((B, 2 * H, 2 * W, w.shape[1]), np.float32)
    x2 = x.reshape(-1, C)
    for dy in range(2):
        for dx in range(2):
            out[:, dy::2, dx::2, :] = (x2 @ w[:, :, dy, dx]).reshape(B, H, W, -1)
    return out + b


def _upconv_bwd(x, w, gout):
    B, H, W, C = x.shape
    x2 = x.reshape(-1, C)
    gw = np.zeros_like(w)
    gx = np.zeros((B * H * W, C), np.float32)
    for dy in range(2):
        for dx in range(2):
            g2 = gout[:, dy::2, dx::2, :].reshape(-1, w.shape[1])
            gw[:, :, dy, dx] = x2.T @ g2
            gx += g2 @ w[:, :, dy, dx].T
    return gx.reshape(x.shape), gw, gout.reshape(-1, w.shape[1]).sum(0)


def depth_of(p, root='out_net'):
    return 1 + max(int(k.split('.')[2]) for k in p if k.startswith(root + '.down_path.'))


def forward(p, x, training, root='out_net'):
    """x: (B, H, W, C) -> ((B, H', W', n_classes), tape).  Mutates the running statistics in ``p`` when training."""
    depth = depth_of(p, root)
    tape = {'down': [], 'pool': [], 'up': []}
    bridges = []
    for i in range(depth):
        x, rec = _block_fwd(p, f'{root}.down_path.{i}.block', x, training)
        tape['down'].append(rec)
        if i != depth - 1:
            bridges.append(x)
            x, prec = _pool_fwd(x)
            tape['pool'].append(prec)
    for i in range(depth - 1):
        pre = f'{root}.up_path.{i}'
        up = _upconv_fwd(x, p[pre + '.up.weight'], p[pre + '.up.bias'])
        bridge = bridges[-i - 1]
        th, tw = up.shape[1:3]
        dy, dx = (bridge.shape[1] - th) // 2, (bridge.shape[2] - tw) // 2          # center_crop, unet.py:118-124
        cat = np.concatenate([up, bridge[:, dy:dy + th, dx:dx + tw]], -1)
        y, rec = _block_fwd(p, pre + '.conv_block.block', cat, training)
        tape['up'].append({'x': x, 'block': rec, 'crop': (dy, dx, th, tw), 'bridge_shape': bridge.shape, 'c_up': up.shape[-1]})
        x = y
    tape['last_in'] = x
    return conv(x, p[root + '.last.weight'], p[root + '.last.bias'], 0), tape


def backward(p, tape, gout, training, root='out_net'):
    """gout: (B, H', W', n_classes) -> (gradient of the U-Net input, {parameter name: gradient})."""
    depth = depth_of(p, root)
    grads = {}
    g, gw, gb = conv_bwd(tape['last_in'], p[root + '.last.weight'], gout, 0)
    grads[root + '.last.weight'], grads[root + '.last.bias'] = gw, gb
    g_bridges = [None] * (depth - 1)
    for i in reversed(range(depth - 1)):
        rec = tape['up'][i]
        pre = f'{root}.up_path.{i}'
        gcat, d = _block_bwd(p, rec['block'], g, training)
        grads.update(d)
        c = rec['c_up']
        dy, dx, th, tw = rec['crop']
        gb_ = np.zeros(rec['bridge_shape'], np.float32)
        gb_[:, dy:dy + th, dx:dx + tw] = gcat[..., c:]
        g_bridges[depth - 2 - i] = gb_                                  # bridges[-i - 1] = bridges[depth - 2 - i]
        g, gw, gbias = _upconv_bwd(rec['x'], p[pre + '.up.weight'], gcat[..., :c])
        grads[pre + '.up.weight'], grads[pre + '.up.bias'] = gw, gbias
    for i in reversed(range(depth)):
        if i != depth - 1:
            g = _pool_bwd(tape['pool'][i], g) + g_bridges[i]
        g, d = _block_bwd(p, tape['down'][i], g, training)
        grads.update(d)
    return g, grads
