"""Oracle (test infrastructure): the FeedForward disparity network in numpy.

Restates /root/reference/mmlf/model/feed_forward.py (topology :25-187, forward
:206-305) with the conv / BatchNorm / ReLU arithmetic that the reference takes
from PyTorch (torch.nn.Conv2d, BatchNorm2d; unpinned dependency, installed
torch 2.11) written out as plain matmuls, plus a hand-derived backward pass so
that gradients can be checked without autograd.

``quant='bf16'`` emulates the storage precision of the CUDA path (bf16
activations and weights, fp32 accumulation and epilogue math) so the kernels
can be compared tightly; ``quant=None`` is the fp32 restatement that is pinned
against the reference's own outputs (tests/golden).
"""
import numpy as np


# ----------------------------------------------------------------------------
# small numeric helpers
# ----------------------------------------------------------------------------
def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what cvt.rn.bf16.f32 does)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(np.float32).reshape(x.shape)
    return np.where(np.isnan(x), x, out)


def fp16_round(x):
    """fp32 -> fp16 (round-to-nearest-even, saturating at +-65504 like the kernels' epilogue) -> fp32."""
    x = np.clip(np.asarray(x, dtype=np.float32), -65504.0, 65504.0)
    return x.astype(np.float16).astype(np.float32)


def torch_linspace_f32(start, end, steps):
    """``torch.linspace(start, end, steps)`` for float32 as ATen's CPU kernel
    computes it (RangeFactories: step in float32; first half ``start + step*i``,
    second half ``end - step*(steps-1-i)``, each a fused multiply-add, i.e. one
    rounding).  Used by class_to_reg / reg_to_class / mpi_to_weights
    (utils/dl.py:126,151,177), always on the CPU (``.to(device)`` afterwards)."""
    start32, end32 = np.float32(start), np.float32(end)
    step = np.float64(np.float32((end32 - start32) / np.float32(steps - 1)))
    i = np.arange(steps)
    half = steps // 2
    up = (np.float64(start32) + step * i).astype(np.float32)
    down = (np.float64(end32) - step * (steps - 1 - i)).astype(np.float32)
    return np.where(i < half, up, down).astype(np.float32)


def np_linspace_f32(start, end, steps):
    """``torch.from_numpy(np.linspace(...))`` assigned into a float32 tensor
    (feed_forward.py:287-288, 298-299; ensamble.py:91-92)."""
    return np.linspace(start, end, steps).astype(np.float32)


def laplacian(x, mu, b):
    """feed_forward.py:9-12 with x (steps,), mu/b (B,H,W) -> (B,steps,H,W)."""
    mu = mu[:, None]
    b = b[:, None]
    x = x.reshape(1, -1, 1, 1)
    one = np.float32(1.0)
    two = np.float32(2.0)
    return (one / (two * b) * np.exp(-np.abs(x - mu) / b)).astype(np.float32)


# ----------------------------------------------------------------------------
# layers (NHWC internally)
# ----------------------------------------------------------------------------
def conv2x2(x, w, b, pad):
    """nn.Conv2d(cin, cout, k, padding=pad) on NHWC x (feed_forward.py:123,125); k = w.shape[2] (2 in the published
    model; odd --model_ksize values use the same code).  w: (cout, cin, k, k), b: (cout,)."""
    k = w.shape[2]
    if pad:
        x = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)))
    B, H, W, C = x.shape
    Ho, Wo = H - k + 1, W - k + 1
    out = np.zeros((B * Ho * Wo, w.shape[0]), np.float32)
    for dy in range(k):
        for dx in range(k):
            a = x[:, dy:dy + Ho, dx:dx + Wo, :].reshape(-1, C)
            out += a @ w[:, :, dy, dx].T
    out += b
    return out.reshape(B, Ho, Wo, -1)


def conv2x2_bwd(x, w, gout, pad):
    """Gradients of conv2x2: returns (gx, gw, gb)."""
    k = w.shape[2]
    xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0))) if pad else x
    B, H, W, C = xp.shape
    Ho, Wo = H - k + 1, W - k + 1
    g2 = gout.reshape(-1, gout.shape[-1])
    gw = np.zeros_like(w)
    gxp = np.zeros_like(xp)
    for dy in range(k):
        for dx in range(k):
            a = xp[:, dy:dy + Ho, dx:dx + Wo, :].reshape(-1, C)
            gw[:, :, dy, dx] = g2.T @ a
            gxp[:, dy:dy + Ho, dx:dx + Wo, :] += (g2 @ w[:, :, dy, dx]).reshape(B, Ho, Wo, C)
    gx = gxp[:, pad:H - pad, pad:W - pad, :] if pad else gxp
    return gx, gw, g2.sum(0)


class FeedForwardOracle:
    """Numpy twin of ``FeedForward`` (feed_forward.py:15).  Parameters are taken
    from a reference ``state_dict`` (Appendix B of SURVEY.md) as numpy arrays."""

    def __init__(self, state, model_cross=False, model_uncert=False, model_discrete=False,
                 model_views=9, model_batchnorm_momentum=0.1, val_disp_min=-3.5,
                 val_disp_max=3.5, quant=None, eps=1e-5):
        self.p = {k: np.array(v, copy=True) for k, v in state.items()}
        self.cross, self.uncert, self.discrete = model_cross, model_uncert, model_discrete
        self.momentum, self.eps, self.quant = model_batchnorm_momentum, eps, quant
        self.disp_min, self.disp_max = val_disp_min, val_disp_max
        self.steps = (2 if model_cross else 4) * model_views * 3      # feed_forward.py:81-84
        self.has_bn = 'in_net_hv.0.3.running_mean' in self.p
        self.in_blocks = 1 + max(int(k.split('.')[1]) for k in self.p if k.startswith('in_net_hv.'))
        self.unet = any(k.startswith('out_net.down_path.') for k in self.p)          # --model_unet (feed_forward.py:99-100)
        self.out_blocks = 0 if self.unet else 1 + max(int(k.split('.')[1]) for k in self.p if k.startswith('out_net.'))
        # kernel size from the weights; paddings as feed_forward.py:86-92 (even k: k//2 then k//2 - 1, odd k: k//2 twice)
        k = self.p['in_net_hv.0.0.weight'].shape[2]
        self.ksize, self.pad1, self.pad2 = k, k // 2, (k // 2 if k % 2 else k // 2 - 1)
        self.training = False

    # -- precision emulation ---------------------------------------------------
    # quant=None: fp32 restatement.  'fp16': what the CUDA path does by default -- forward activations and
    # weights stored in fp16 (saturating), gradients and the dgrad weight operand in bf16.  'bf16': everything bf16.
    def _q(self, x):
        if self.quant == 'fp16':
            return fp16_round(x)
        return bf16_round(x) if self.quant == 'bf16' else x

    def _qg(self, x):
        return bf16_round(x) if self.quant else x

    def _w(self, name, grad=False):
        w = self.p[name]
        if not self.quant:
            return w
        return bf16_round(w) if (grad or self.quant == 'bf16') else fp16_round(w)

    # -- one block: conv(k2,p1) -> ReLU -> conv(k2,p0) [-> BN -> ReLU]  (feed_forward.py:122-137)
    def _block_fwd(self, prefix, x, bn, tape, head_fp32=False, relu_out=True):
        w1, b1 = self._w(prefix + '.0.weight'), self.p[prefix + '.0.bias']
        w2, b2 = self._w(prefix + '.2.weight'), self.p[prefix + '.2.bias']
        a1 = np.maximum(conv2x2(x, w1, b1, self.pad1), 0)
        a1 = a1 if head_fp32 else self._q(a1)
        if head_fp32 and self.quant:
            w2 = self.p[prefix + '.2.weight']          # tiny head conv2 runs in fp32 on CUDA cores
        z = conv2x2(a1, w2, b2, self.pad2)
        rec = {'prefix': prefix, 'x': x, 'a1': a1, 'bn': bn, 'relu_out': relu_out}
        if not bn:
            if relu_out:                       # --model_no_batchnorm: index 3 is the ReLU (feed_forward.py:132-135)
                z = self._q(np.maximum(z, 0))
                rec['y'] = z
            tape.append(rec)
            return z
        g, be = self.p[prefix + '.3.weight'], self.p[prefix + '.3.bias']
        if self.training:
            n = z.shape[0] * z.shape[1] * z.shape[2]
            zf = z.reshape(-1, z.shape[-1]).astype(np.float64)
            mean = zf.mean(0)
            var = zf.var(0)                                   # biased, used to normalise
            rm, rv = self.p[prefix + '.3.running_mean'], self.p[prefix + '.3.running_var']
            m = self.momentum
            self.p[prefix + '.3.running_mean'] = ((1 - m) * rm + m * mean).astype(np.float32)
            self.p[prefix + '.3.running_var'] = ((1 - m) * rv + m * var * n / (n - 1)).astype(np.float32)
            self.p[prefix + '.3.num_batches_tracked'] = self.p[prefix + '.3.num_batches_tracked'] + 1
            mean, var = mean.astype(np.float32), var.astype(np.float32)
            zq = self._q(z)                                   # CUDA path stores z as bf16, stats from fp32
        else:
            mean, var = self.p[prefix + '.3.running_mean'], self.p[prefix + '.3.running_var']
            # eval: BN folded into the conv epilogue (z never stored); a differentiable eval-mode forward
            # (--train_eval_mode) keeps z for the backward pass in the activation format like the training path
            zq = self._q(z) if getattr(self, '_keep_z', False) else z
        invstd = (1.0 / np.sqrt(var.astype(np.float64) + self.eps)).astype(np.float32)
        xhat = (zq - mean) * invstd
        y = self._q(np.maximum(xhat * g + be, 0))
        rec.update(xhat=xhat, invstd=invstd, y=y)
        tape.append(rec)
        return y

    def _block_bwd(self, rec, gy, need_gx=True):
        prefix = rec['prefix']
        grads = {}
        if rec['bn']:
            g = self.p[prefix + '.3.weight']
            gy = gy * (rec['y'] > 0)
            gy2 = gy.reshape(-1, gy.shape[-1])
            xh2 = rec['xhat'].reshape(-1, gy.shape[-1])
            grads[prefix + '.3.weight'] = (gy2 * xh2).sum(0)
            grads[prefix + '.3.bias'] = gy2.sum(0)
            if self.training:
                n = gy2.shape[0]
                gz = (g * rec['invstd']) * (gy - gy2.mean(0) - rec['xhat'] * (gy2 * xh2).sum(0) / n)
            else:
                gz = gy * (g * rec['invstd'])
            gz = self._qg(gz)
        else:
            gz = self._qg(gy * (rec['y'] > 0)) if rec['relu_out'] else gy
        w1, w2 = self._w(prefix + '.0.weight', grad=True), self._w(prefix + '.2.weight', grad=True)
        # the weight-gradient GEMM reads its activation operand converted to the gradient format (bf16)
        ga1, gw2, gb2 = conv2x2_bwd(self._qg(rec['a1']), w2, gz, self.pad2)
        ga1 = self._qg(ga1 * (rec['a1'] > 0))
        gx, gw1, gb1 = conv2x2_bwd(self._qg(rec['x']), w1, ga1, self.pad1)
        grads[prefix + '.0.weight'], grads[prefix + '.0.bias'] = gw1, gb1
        grads[prefix + '.2.weight'], grads[prefix + '.2.bias'] = gw2, gb2
        return (self._qg(gx) if need_gx else None), grads

    def _in_net(self, name, x, tape):
        for k in range(self.in_blocks):
            x = self._block_fwd(f'{name}.{k}', x, self.has_bn, tape)
        return x

    # -- forward (feed_forward.py:206-305) ---------------------------------------
    def forward(self, h_views, v_views, i_views=None, d_views=None, keep_tape=False):
        b, n, c, h, w = h_views.shape
        self._keep_z = keep_tape
        nhwc = lambda t: self._q(np.ascontiguousarray(  # noqa: E731
            t.reshape(b, n * c, h, w).transpose(0, 2, 3, 1)))
        tapes = {}
        # h: permute(0,1,3,2) -> in_net_hv -> permute back  (feed_forward.py:236-241)
        tapes['h'] = []
        hf = self._in_net('in_net_hv', nhwc(h_views).transpose(0, 2, 1, 3), tapes['h']).transpose(0, 2, 1, 3)
        tapes['v'] = []
        vf = self._in_net('in_net_hv', nhwc(v_views), tapes['v'])
        feats = [hf, vf]
        if not self.cross:
            # i: permute, flip(-1), in_net_id, flip(-1), permute  (feed_forward.py:248-256)
            tapes['i'] = []
            xi = nhwc(i_views).transpose(0, 2, 1, 3)[:, :, ::-1, :]
            fi = self._in_net('in_net_id', np.ascontiguousarray(xi), tapes['i'])
            feats.append(fi[:, :, ::-1, :].transpose(0, 2, 1, 3))
            tapes['d'] = []
            feats.append(self._in_net('in_net_id', nhwc(d_views), tapes['d']))
        x = np.concatenate(feats, -1)                                  # feed_forward.py:263-267
        tapes['o'] = []
        if self.unet:
            from . import unet
            out, tapes['unet'] = unet.forward(self.p, x, self.training)
        else:
            for k in range(self.out_blocks - 1):
                x = self._block_fwd(f'out_net.{k}', x, self.has_bn, tapes['o'])
            small_head = not self.discrete
            out = self._block_fwd(f'out_net.{self.out_blocks - 1}', x, False, tapes['o'],
                                  head_fp32=small_head, relu_out=False)
        output = np.ascontiguousarray(out.transpose(0, 3, 1, 2))        # NCHW
        if keep_tape:
            self._tapes = tapes
        res = {'output': output, 'mean': output[:, 0], 'logvar': None, 'scores': None,
               'one_hot': None, 'posterior': None}
        if self.discrete:                                              # feed_forward.py:276-290
            s = output
            res['scores'] = s
            res['one_hot'] = (s.max(1, keepdims=True) == s).astype(np.float32)
            e = np.exp(s)
            res['posterior'] = e / e.sum(1, keepdims=True, dtype=np.float32)
            bins_t = torch_linspace_f32(self.disp_min, self.disp_max, self.steps).reshape(1, -1, 1, 1)
            res['mean'] = (bins_t * res['one_hot']).sum(1, dtype=np.float32)
            bins_n = np_linspace_f32(self.disp_min, self.disp_max, self.steps).reshape(1, -1, 1, 1)
            lv = (bins_n - res['mean'][:, None]) ** np.float32(2.0) * res['posterior']
            with np.errstate(divide='ignore'):
                res['logvar'] = np.log(lv.sum(1, dtype=np.float32))
        if self.uncert:                                                # feed_forward.py:292-302
            res['logvar'] = output[:, 1]
            x = np_linspace_f32(self.disp_min, self.disp_max, self.steps)
            res['posterior'] = laplacian(x, res['mean'], np.exp(res['logvar']))
        return res

    # -- backward: d loss / d output (B, OC, H, W) -> parameter gradients ---------
    def backward(self, g_output):
        tapes = self._tapes
        grads = {}

        def acc(d):
            for k, v in d.items():
                grads[k] = grads[k] + v if k in grads else v

        g = np.ascontiguousarray(g_output.transpose(0, 2, 3, 1)).astype(np.float32)
        if self.unet:
            from . import unet
            g, d = unet.backward(self.p, tapes['unet'], g, self.training)
            acc(d)
        for rec in reversed(tapes['o']):
            g, d = self._block_bwd(rec, g)
            acc(d)
        c = g.shape[-1] // (2 if self.cross else 4)
        parts = [g[..., k * c:(k + 1) * c] for k in range(g.shape[-1] // c)]
        # undo the layout plumbing of forward()
        gin = {'h': parts[0].transpose(0, 2, 1, 3), 'v': parts[1]}
        if not self.cross:
            gin['i'] = np.ascontiguousarray(parts[2].transpose(0, 2, 1, 3)[:, :, ::-1, :])
            gin['d'] = parts[3]
        for key, gg in gin.items():
            gg = np.ascontiguousarray(gg)
            recs = tapes[key]
            for j, rec in enumerate(reversed(recs)):
                gg, d = self._block_bwd(rec, gg, need_gx=(j != len(recs) - 1))
                acc(d)
        return grads
