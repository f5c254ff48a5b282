"""Execution plan for the topologies outside the published 2x2-kernel network: odd
``--model_ksize`` (symmetric padding k // 2 for both convs of a block, /root/reference/mmlf/model/feed_forward.py:86-92)
and the ``--model_unet`` out-net (feed_forward.py:99-100, 189-204; /root/reference/mmlf/model/unet.py:8-132).

Everything runs on the float32 CUDA-core layer kernels of csrc/generic.cu over dense channel-last tensors -- forward and
a hand-written backward, same interface as :class:`mmlf_b200.engine.Engine` (``forward`` -> output + tape, ``backward`` ->
gradients in one flat buffer), so FeedForward, the autograd node, TrainStep and the inference graphs work unchanged.  The
stream plumbing of feed_forward.py:236-256 (transposes / flips of the h and i stacks) is folded into the weight packing
(``mmlf_g_pack_weight`` ``spatial``), never applied to data.  These rows are about coverage and fp32-exact parity; the
tensor-core path is the published topology's.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import call


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Dense:
    """A channel-last float32 activation: B x H x W pixels, C channels starting at column c0 of the 2-D buffer `buf`."""

    def __init__(self, buf, B, H, W, C_, c0=0):
        self.buf, self.B, self.H, self.W, self.C, self.c0 = buf, B, H, W, C_, c0
        self.ld = buf.shape[1]
        self.rows = B * H * W

    @property
    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 4 * self.c0)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class GenericEngine:
    def __init__(self, module):
        self.m = module
        self.k = module.ksize
        self.p1, self.p2 = module.padding1, module.padding2
        self.unet = bool(getattr(module, 'unet', False))
        self.has_bn = not module.no_batchnorm
        self.cross = module.cross
        self.chs = module.chs
        self.oc = module.out_chs
        self.stream_defs = [('h', 'in_net_hv', 1), ('v', 'in_net_hv', 0)]
        if not self.cross:
            self.stream_defs += [('i', 'in_net_id', 2), ('d', 'in_net_id', 0)]
        self._param_cache = self._buffer_cache = self._module_cache = None
        self._glayout = None
        self._gflat = None
        self.dev = None

    # ------------------------------------------------------------------ parameters
    def _params(self):
        if self._param_cache is None:
            self._param_cache = dict(self.m.named_parameters())
        return self._param_cache

    def _buffers(self):
        if self._buffer_cache is None:
            self._buffer_cache = dict(self.m.named_buffers())
        return self._buffer_cache

    def _modules(self):
        if self._module_cache is None:
            self._module_cache = dict(self.m.named_modules())
        return self._module_cache

    def grad_layout(self):
        if self._glayout is None:
            off, lay = 0, {}
            for name, p in self._params().items():
                lay[name] = (off, p.numel(), tuple(p.shape))
                off += p.numel()
            self._glayout = (lay, off)
        return self._glayout

    # ------------------------------------------------------------------ primitive layers
    def _new(self, B, H, W, C_):
        return Dense(torch.empty((B * H * W, C_), dtype=torch.float32, device=self.dev), B, H, W, C_)

    def _zeros64(self, n):
        t = torch.empty(n, dtype=torch.float64, device=self.dev)
        call('mmlf_zero', _p(t), n * 8, _st())
        return t

    def _packed(self, name, mode, spatial=0):
        w = self._params()[name + '.weight'].detach()
        if mode <= 1:
            cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        else:                                               # ConvTranspose2d weight (cin, cout, 2, 2)
            cin, cout, k = w.shape[0], w.shape[1], 2
        out = torch.empty(w.numel(), dtype=torch.float32, device=self.dev)
        call('mmlf_g_pack_weight', _p(w), cout, cin, k, spatial, mode, _p(out), _st())
        return out

    def _conv(self, x, name, pad, relu, spatial=0, out=None):
        w = self._params()[name + '.weight']
        b = self._params()[name + '.bias'].detach()
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        assert cin == x.C, (name, cin, x.C)
        Ho, Wo = x.H + 2 * pad - k + 1, x.W + 2 * pad - k + 1
        y = out if out is not None else self._new(x.B, Ho, Wo, cout)
        call('mmlf_g_conv', x.ptr, x.ld, _p(self._packed(name, 0, spatial)), _p(b), x.B, x.H, x.W, cin, cout, k, pad,
             1 if relu else 0, y.ptr, y.ld, _st())
        return y

    def _conv_bwd(self, x, name, pad, gz, spatial, need_gx, gp):
        """Gradients of y = conv(x): weight + bias gradients are ADDED into the flat buffer (gp: name -> pointer),
        returns dL/dx (or None)."""
        w = self._params()[name + '.weight']
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        call('mmlf_g_conv_wgrad', x.ptr, x.ld, gz.ptr, gz.ld, x.B, x.H, x.W, cin, cout, k, pad, spatial, 0,
             gp(name + '.weight'), _st())
        call('mmlf_g_colsum', gz.ptr, gz.ld, cout, gz.rows, gp(name + '.bias'), _st())
        if not need_gx:
            return None
        gx = self._new(x.B, x.H, x.W, cin)
        call('mmlf_g_conv', gz.ptr, gz.ld, _p(self._packed(name, 1, spatial)), C.c_void_p(0), gz.B, gz.H, gz.W, cout, cin, k,
             k - 1 - pad, 0, gx.ptr, gx.ld, _st())
        return gx

    def _bn(self, x, name, training, relu, out=None):
        """BatchNorm2d `name` on x (+ optional ReLU) -> (y, record for the backward pass)."""
        params, bufs = self._params(), self._buffers()
        mod = self._modules()[name]
        gamma, beta = params[name + '.weight'].detach(), params[name + '.bias'].detach()
        rmean, rvar = bufs[name + '.running_mean'], bufs[name + '.running_var']
        Cc = x.C
        consts = torch.empty((4, Cc), dtype=torch.float32, device=self.dev)
        scale, shift, mean, invstd = consts[0], consts[1], consts[2], consts[3]
        if training:
            sums = self._zeros64(2 * Cc)
            call('mmlf_g_bn_stats', x.ptr, x.ld, Cc, x.rows, _p(sums), _st())
            nbt = bufs.get(name + '.num_batches_tracked')
            call('mmlf_bn_finalize', _p(sums), Cc, Cc, x.rows, _p(gamma), _p(beta), _p(rmean), _p(rvar), _p(nbt),
                 float(mod.momentum), float(mod.eps), _p(scale), _p(shift), _p(mean), _p(invstd), _st())
            torch.autograd.graph.increment_version([t for t in (rmean, rvar, nbt) if t is not None])
        else:
            call('mmlf_bn_fold_eval', Cc, Cc, _p(gamma), _p(beta), _p(rmean), _p(rvar), C.c_void_p(0), float(mod.eps),
                 _p(scale), _p(shift), _st())
            mean.copy_(rmean)
            invstd.copy_(torch.rsqrt(rvar + float(mod.eps)))
        y = out if out is not None else self._new(x.B, x.H, x.W, Cc)
        call('mmlf_g_affine', x.ptr, x.ld, _p(scale), _p(shift), Cc, x.rows, 1 if relu else 0, y.ptr, y.ld, _st())
        return y, dict(name=name, x=x, y=y if relu else None, mean=mean, invstd=invstd, train=training)

    def _bn_bwd(self, rec, gy, gp):
        x = rec['x']
        gamma = self._params()[rec['name'] + '.weight'].detach()
        gx = self._new(x.B, x.H, x.W, x.C)
        gate = rec['y']
        sums = self._zeros64(2 * x.C)
        call('mmlf_g_bn_bwd', gy.ptr, gy.ld, x.ptr, x.ld, gate.ptr if gate is not None else C.c_void_p(0),
             gate.ld if gate is not None else 0, _p(gamma), _p(rec['mean']), _p(rec['invstd']), _p(sums), x.rows,
             1 if rec['train'] else 0, x.C, x.rows, gx.ptr, gx.ld, gp(rec['name'] + '.weight'), gp(rec['name'] + '.bias'),
             _st())
        return gx

    def _relu_bwd(self, gy, y):
        out = self._new(y.B, y.H, y.W, y.C)
        call('mmlf_g_relu_bwd', gy.ptr, gy.ld, y.ptr, y.ld, y.C, y.rows, out.ptr, out.ld, _st())
        return out

    # ------------------------------------------------------------------ blocks
    def _block_fwd(self, prefix, x, training, spatial, out_bn_relu=True, out=None):
        """feed_forward.py:106-137: conv(k, p1) -> ReLU -> conv(k, p2) [-> BatchNorm] [-> ReLU]."""
        a1 = self._conv(x, prefix + '.0', self.p1, True, spatial)
        rec = dict(kind='block', prefix=prefix, x=x, a1=a1, spatial=spatial, bn=None, y=None)
        if not out_bn_relu:
            z = self._conv(a1, prefix + '.2', self.p2, False, spatial, out=out)
            return z, rec
        if not self.has_bn:
            y = self._conv(a1, prefix + '.2', self.p2, True, spatial, out=out)
            rec['y'] = y
            return y, rec
        z = self._conv(a1, prefix + '.2', self.p2, False, spatial)
        y, rec['bn'] = self._bn(z, prefix + '.3', training, True, out=out)
        return y, rec

    def _block_bwd(self, rec, gy, need_gx, gp):
        prefix, sp = rec['prefix'], rec['spatial']
        if rec['bn'] is not None:
            gz = self._bn_bwd(rec['bn'], gy, gp)
        elif rec['y'] is not None:
            gz = self._relu_bwd(gy, rec['y'])
        else:
            gz = gy
        ga1 = self._conv_bwd(rec['a1'], prefix + '.2', self.p2, gz, sp, True, gp)
        ga1 = self._relu_bwd(ga1, rec['a1'])
        return self._conv_bwd(rec['x'], prefix + '.0', self.p1, ga1, sp, need_gx, gp)

    def _ublock_fwd(self, prefix, x, training):
        """unet.py:80-101: [conv3x3 p1 -> ReLU -> BatchNorm] x 2."""
        a1 = self._conv(x, prefix + '.0', 1, True)
        y1, bn1 = self._bn(a1, prefix + '.2', training, False)
        a2 = self._conv(y1, prefix + '.3', 1, True)
        y2, bn2 = self._bn(a2, prefix + '.5', training, False)
        return y2, dict(prefix=prefix, x=x, a1=a1, bn1=bn1, y1=y1, a2=a2, bn2=bn2)

    def _ublock_bwd(self, rec, gy, gp, need_gx=True):
        prefix = rec['prefix']
        ga2 = self._relu_bwd(self._bn_bwd(rec['bn2'], gy, gp), rec['a2'])
        gy1 = self._conv_bwd(rec['y1'], prefix + '.3', 1, ga2, 0, True, gp)
        ga1 = self._relu_bwd(self._bn_bwd(rec['bn1'], gy1, gp), rec['a1'])
        return self._conv_bwd(rec['x'], prefix + '.0', 1, ga1, 0, need_gx, gp)

    # ------------------------------------------------------------------ U-Net (unet.py:62-77)
    def _unet_fwd(self, x, training):
        un = self.m.out_net
        depth = un.depth
        tape = dict(down=[], pool=[], up=[])
        bridges = []
        for i in range(depth):
            x, rec = self._ublock_fwd(f'out_net.down_path.{i}.block', x, training)
            tape['down'].append(rec)
            if i != depth - 1:
                bridges.append(x)
                if x.H < 2 or x.W < 2:
                    raise RuntimeError(f'--model_unet: the input is too small for {depth - 1} poolings')
                y = self._new(x.B, x.H // 2, x.W // 2, x.C)
                idx = torch.empty(y.rows * y.C, dtype=torch.uint8, device=self.dev)
                call('mmlf_g_maxpool2', x.ptr, x.B, x.H, x.W, x.C, y.ptr, _p(idx), _st())
                tape['pool'].append(dict(idx=idx, shape=(x.B, x.H, x.W, x.C)))
                x = y
        for i in range(depth - 1):
            pre = f'out_net.up_path.{i}'
            w = self._params()[pre + '.up.weight']
            cin, cu = w.shape[0], w.shape[1]
            bias4 = self._params()[pre + '.up.bias'].detach().repeat(4)
            y4 = torch.empty((x.rows, 4 * cu), dtype=torch.float32, device=self.dev)
            call('mmlf_g_conv', x.ptr, x.ld, _p(self._packed(pre + '.up', 2)), _p(bias4), x.B, x.H, x.W, cin, 4 * cu, 1, 0, 0,
                 _p(y4), 4 * cu, _st())
            bridge = bridges[-i - 1]
            th, tw = 2 * x.H, 2 * x.W
            cat = self._new(x.B, th, tw, cu + bridge.C)
            call('mmlf_g_depth_to_space', _p(y4), cat.ptr, cat.ld, 0, x.B, x.H, x.W, cu, 0, _st())
            dy, dx = (bridge.H - th) // 2, (bridge.W - tw) // 2                  # center_crop, unet.py:118-124
            call('mmlf_g_copy_window', bridge.ptr, bridge.H, bridge.W, bridge.ld, 0, dy, dx, cat.ptr, th, tw, cat.ld, cu, 0, 0,
                 x.B, th, tw, bridge.C, 0, _st())
            y, rec = self._ublock_fwd(pre + '.conv_block.block', cat, training)
            tape['up'].append(dict(x=x, block=rec, crop=(dy, dx), bridge=bridge, cu=cu, pre=pre))
            x = y
        tape['last_in'] = x
        return self._conv(x, 'out_net.last', 0, False), tape

    def _unet_bwd(self, tape, g, gp):
        depth = self.m.out_net.depth
        g = self._conv_bwd(tape['last_in'], 'out_net.last', 0, g, 0, True, gp)
        g_bridges = [None] * (depth - 1)
        for i in reversed(range(depth - 1)):
            rec = tape['up'][i]
            pre, x, bridge, cu = rec['pre'], rec['x'], rec['bridge'], rec['cu']
            gcat = self._ublock_bwd(rec['block'], g, gp)
            dy, dx = rec['crop']
            gb = self._new(bridge.B, bridge.H, bridge.W, bridge.C)
            if (gcat.H, gcat.W) != (bridge.H, bridge.W):
                call('mmlf_zero', gb.ptr, gb.rows * gb.C * 4, _st())
            call('mmlf_g_copy_window', gcat.ptr, gcat.H, gcat.W, gcat.ld, cu, 0, 0, gb.ptr, gb.H, gb.W, gb.ld, 0, dy, dx,
                 gcat.B, gcat.H, gcat.W, bridge.C, 0, _st())
            g_bridges[depth - 2 - i] = gb
            # transposed conv backward: space-to-depth of the gradient, then two GEMMs and a column sum
            g4 = torch.empty((x.rows, 4 * cu), dtype=torch.float32, device=self.dev)
            call('mmlf_g_depth_to_space', _p(g4), gcat.ptr, gcat.ld, 0, x.B, x.H, x.W, cu, 1, _st())
            cin = x.C
            call('mmlf_g_conv_wgrad', x.ptr, x.ld, _p(g4), 4 * cu, x.B, x.H, x.W, cin, 4 * cu, 1, 0, 0, 1,
                 gp(pre + '.up.weight'), _st())
            call('mmlf_g_colsum', gcat.ptr, gcat.ld, cu, gcat.rows, gp(pre + '.up.bias'), _st())
            g = self._new(x.B, x.H, x.W, cin)
            call('mmlf_g_conv', _p(g4), 4 * cu, _p(self._packed(pre + '.up', 3)), C.c_void_p(0), x.B, x.H, x.W, 4 * cu, cin, 1,
                 0, 0, g.ptr, g.ld, _st())
        for i in reversed(range(depth)):
            if i != depth - 1:
                B, H, W, Cc = tape['pool'][i]['shape']
                gx = self._new(B, H, W, Cc)
                call('mmlf_g_maxpool2_bwd', g.ptr, _p(tape['pool'][i]['idx']), B, H, W, Cc, gx.ptr, _st())
                gbr = g_bridges[i]
                call('mmlf_g_copy_window', gbr.ptr, H, W, gbr.ld, 0, 0, 0, gx.ptr, H, W, gx.ld, 0, 0, 0, B, H, W, Cc, 1, _st())
                g = gx
            g = self._ublock_bwd(tape['down'][i], g, gp)
        return g

    # ------------------------------------------------------------------ forward / backward of the whole network
    def forward(self, views, training, save, shift_disp=None):
        _lib.require_device()
        for v in views:
            if not (isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
                raise RuntimeError('mmlf_b200: view stacks must be contiguous float32 CUDA tensors; there is no CPU path')
        p0 = next(iter(self._params().values()))
        if p0.device != views[0].device:
            raise RuntimeError(f'mmlf_b200: the model lives on {p0.device} but the inputs on {views[0].device}')
        if getattr(self.m, 'precision', 'fp16') == 'split':
            raise RuntimeError("precision='split' applies to the 2x2 tensor-core path; this topology already runs in float32")
        if shift_disp is not None:                                  # ESE member: bit-exact Shift kernel first
            from . import ops
            full = list(views) + [views[0]] * (4 - len(views))
            views = ops.lf_shift(*full, float(shift_disp))[:len(views)]
        B, n, c3, H, W = views[0].shape
        self.dev = views[0].device
        bn_train = training and self.has_bn
        width = len(self.stream_defs) * self.chs
        feats = self._new(B, H, W, width)
        tape = dict(streams={}, out=[], bn_train=bn_train, geo=(B, H, W))
        blocks = self.m.n_in_blocks
        for si, (key, net, spatial) in enumerate(self.stream_defs):
            x = self._new(B, H, W, n * c3)
            call('mmlf_g_layout', _p(views[si]), x.ptr, x.ld, B, n * c3, H, W, 1, _st())
            recs = []
            for kb in range(blocks):
                out = Dense(feats.buf, B, H, W, self.chs, si * self.chs) if kb == blocks - 1 else None
                x, rec = self._block_fwd(f'{net}.{kb}', x, bn_train, spatial, out=out)
                recs.append(rec)
            tape['streams'][key] = recs
        x = feats
        if self.unet:
            y, tape['unet'] = self._unet_fwd(x, bn_train)
        else:
            nb = self.m.n_out_blocks
            for kb in range(nb - 1):
                x, rec = self._block_fwd(f'out_net.{kb}', x, bn_train, 0)
                tape['out'].append(rec)
            y, rec = self._block_fwd(f'out_net.{nb - 1}', x, bn_train, 0, out_bn_relu=False)
            tape['out'].append(rec)
        out = torch.empty((B, y.C, y.H, y.W), dtype=torch.float32, device=self.dev)
        call('mmlf_g_layout', _p(out), y.ptr, y.ld, B, y.C, y.H, y.W, 0, _st())
        return out, (tape if save else None)

    def backward(self, tape, g_out, flat=None):
        layout, total = self.grad_layout()
        dev = g_out.device
        self.dev = dev
        params = self._params()
        if flat is None:
            lo = self._gflat.data_ptr() if self._gflat is not None else 0
            if self._gflat is None or self._gflat.device != dev or \
                    any(p.grad is not None and lo <= p.grad.data_ptr() < lo + 4 * total for p in params.values()):
                self._gflat = torch.empty(total, dtype=torch.float32, device=dev)
            flat = self._gflat
            call('mmlf_zero', _p(flat), total * 4, _st())
        base = flat.data_ptr()

        def gp(name):
            return C.c_void_p(base + 4 * layout[name][0])

        B, Cc, H, W = g_out.shape
        g = self._new(B, H, W, Cc)
        call('mmlf_g_layout', _p(g_out.contiguous()), g.ptr, g.ld, B, Cc, H, W, 1, _st())
        if self.unet:
            g = self._unet_bwd(tape['unet'], g, gp)
        else:
            for rec in reversed(tape['out']):
                g = self._block_bwd(rec, g, True, gp)
        for si, (key, net, spatial) in enumerate(self.stream_defs):
            gs = Dense(g.buf, g.B, g.H, g.W, self.chs, si * self.chs)
            recs = tape['streams'][key]
            for j in reversed(range(len(recs))):
                gs = self._block_bwd(recs[j], gs, j != 0, gp)
        return {name: flat[off:off + n].view(shape) for name, (off, n, shape) in layout.items()}
