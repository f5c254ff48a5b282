"""Execution plan of the FeedForward disparity network on the sm_100a kernels.

Mirrors ``FeedForward.forward`` of the reference (/root/reference/mmlf/model/feed_forward.py:206-305) and its
autograd, but on the bf16 "slot" layout (include/mmlf_b200.h): every conv is one launch of the tcgen05 implicit
GEMM, BatchNorm/ReLU/heads are fused or bandwidth kernels, and the backward pass is written out by hand
(dgrad = the same conv kernel with rotated weights, wgrad = the MN-major tcgen05 kernel).

PyTorch is used for device memory, streams and autograd plumbing only; no ATen compute op runs on the path.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import BF16, FP16, ConvArgs, call

_TORCH_DT = {BF16: torch.bfloat16, FP16: torch.float16}
GRAD = BF16      # gradients are always bf16 (fp32 exponent range)
SPLIT_WSCALE = 64.0   # split precision: weights are packed as w * 64 so that their fp16 residuals stay normal numbers


def pad16(x):
    return (x + 15) // 16 * 16


def bits_words(n_pad):
    """Words per slot of a ReLU sign-bit array (one bit per channel)."""
    return (n_pad + 31) // 32


class PackJob(C.Structure):
    """Mirror of ``mmlf_pack_job`` (include/mmlf_b200.h): one weight packing of the batched launch."""
    _fields_ = [('w', C.c_void_p), ('out', C.c_void_p), ('bias', C.c_void_p), ('bias_pad', C.c_void_p),
                ('cout', C.c_int), ('cin', C.c_int), ('spatial', C.c_int), ('dgrad', C.c_int), ('in_groups', C.c_int),
                ('group_real', C.c_int), ('group_pad', C.c_int), ('n_pad', C.c_int), ('cin_pad', C.c_int),
                ('dtype', C.c_int), ('split', C.c_int), ('weight_scale', C.c_float)]


class VecJob(C.Structure):
    """Mirror of ``mmlf_vec_job`` (include/mmlf_b200.h): one short vector update of the end-of-backward launch."""
    _fields_ = [('src', C.c_void_p), ('src2', C.c_void_p), ('dst', C.c_void_p), ('n', C.c_int32), ('src_f64', C.c_int32),
                ('accumulate', C.c_int32), ('pad_', C.c_int32)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _zero(t):
    """cudaMemsetAsync on the current stream (no framework fill kernel)."""
    call('mmlf_zero', _ptr(t), t.numel() * t.element_size(), _stream())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Geometry:
    def __init__(self, B, H, W):
        self.B, self.H, self.W = B, H, W
        self.Hp, self.Wp = H + 1, W + 1
        self.n_slots = B * self.Hp * self.Wp
        self.count = B * H * W


class ConvSpec:
    """One nn.Conv2d of the reference with its padded GEMM shapes and packed operands."""

    def __init__(self, name, cout, cin, ctype, spatial=0, groups=1, group_real=None, group_pad=None):
        self.name = name                      # state_dict prefix, e.g. 'out_net.0.0'
        self.cout, self.cin, self.type, self.spatial = cout, cin, ctype, spatial
        self.groups = groups
        self.group_real = group_real if group_real is not None else cin
        self.group_pad = group_pad if group_pad is not None else pad16(cin)
        self.cin_pad = groups * self.group_pad if groups > 1 else pad16(cin)
        self.n_pad = pad16(cout)
        self.w_fwd = None                     # bf16 [n_pad][4*kc*64]
        self.w_split = None                   # fp16 [n_pad][4*3*kc*64] (split-precision inference)
        self.unscale = None                   # f32 [n_pad] = 1 / SPLIT_WSCALE
        self.w_dgrad = None                   # bf16 [cin_pad][4*kc'*64]
        self.bias_pad = None                  # f32 [n_pad]

    @property
    def kc(self):
        return (self.cin_pad + 63) // 64


class Engine:
    """Owns packed parameters and scratch buffers for one FeedForward module."""

    def __init__(self, module):
        self.m = module
        self.chs = module.chs
        self.views = module.views
        self.cross = module.cross
        self.has_bn = not module.no_batchnorm
        self.n_streams = 2 if self.cross else 4
        self.in_blocks = module.n_in_blocks
        self.out_blocks = module.n_out_blocks
        self.oc = module.out_chs
        self.small_head = self.oc <= 2
        self.cp = pad16(self.chs)                       # per-stream feature pitch (70 -> 80)
        self.feat_ld = self.n_streams * self.cp         # concatenated feature buffer (280 -> 320)
        self.width = self.n_streams * self.chs          # out-net width (280)
        self.wp = pad16(self.width)                     # 288
        assert self.feat_ld <= 320 and self.wp <= 320, 'model_chs too large for the 320-column TMEM plan'
        self._pack_version = None
        self._pack_jobs = {}          # key -> device job table; never dropped: captured graphs hold their addresses
        self._fold_cache = {}
        self._param_cache = None
        self._buffer_cache = None
        self._scratch_cache = {}
        self._gflat = None
        self._glayout = None
        self._vec_jobs = {}           # key -> device job table (kept for the same reason)
        self._ws = None
        self._pinned_tables = set()
        self._side = None
        # BatchNorm-backward statistics out of the data-gradient epilogue: implemented and tested, but opt-in
        # (MMLF_BN_FUSE=1).  Measured on B200: the wide data gradient goes from 0.293 to 0.382 ms (the epilogue becomes the
        # bound) against 0.141 ms for the standalone reduction -- 207.1 vs 207.6 ms per B = 512 step, i.e. nothing, while
        # the conv kernel's own roofline fraction drops from 0.69 to 0.65.
        self.fuse_bn_bwd = os.environ.get('MMLF_BN_FUSE', '0') == '1'
        self.overlap_wgrad = os.environ.get('MMLF_OVERLAP_WGRAD', '0') == '1'
        # MMLF_SINGLE_ACT=1: keep ONE copy of every activation; the weight-gradient GEMM then reads the fp16 activations
        # against the bf16 gradients by converting its activation boxes in shared memory (conv2x2_wgrad.cu).  Saves 30 % of
        # the activation memory, but measured on B200 it is a wash to a loss in time: bn_apply_relu 183 -> 138 us and the
        # first conv of a block ~ -45 us per out-net block, against +33..40 us on EACH of the block's two weight-gradient
        # launches (351 -> 385 us: the extra shared-memory traffic of the in-place conversion competes with the MMA operand
        # fetch); B = 512 step 209.8 -> 219.5 ms.  Default: the round-1 scheme, a bf16 twin of every activation written by
        # the kernel that produces it.
        self.mixed_wgrad = os.environ.get('MMLF_SINGLE_ACT', '0') == '1'
        self._build_specs()

    @property
    def act(self):
        """Storage format of forward activations and weights ('fp16' default, 'bf16' optional, 'split' = fp16 hi + lo)."""
        return BF16 if getattr(self.m, 'precision', 'fp16') == 'bf16' else FP16

    @property
    def split(self):
        """Split-precision inference (``model.precision = 'split'``): every activation and weight is carried as
        fp16 hi + fp16 lo and every product as hi*hi + hi*lo + lo*hi in the fp32 accumulator -- fp32-class results on the
        fp16 tensor cores at three times the MMA work.  Eval / no-grad only."""
        return getattr(self.m, 'precision', 'fp16') == 'split'

    # ------------------------------------------------------------------ static plan
    def _build_specs(self):
        cin0 = self.views * 3
        self.stream_defs = [('h', 'in_net_hv', 1), ('v', 'in_net_hv', 0)]
        if not self.cross:
            self.stream_defs += [('i', 'in_net_id', 2), ('d', 'in_net_id', 0)]
        self.in_specs = {}
        for key, net, spatial in self.stream_defs:
            blocks = []
            for k in range(self.in_blocks):
                cin = cin0 if k == 0 else self.chs
                c1 = ConvSpec(f'{net}.{k}.0', self.chs, cin, 0, spatial)
                c2 = ConvSpec(f'{net}.{k}.2', self.chs, self.chs, 1, spatial)
                blocks.append((c1, c2, f'{net}.{k}.3'))
            self.in_specs[key] = blocks
        self.out_specs = []
        for k in range(self.out_blocks - 1):
            if k == 0:
                c1 = ConvSpec(f'out_net.{k}.0', self.width, self.width, 0, 0, self.n_streams, self.chs, self.cp)
            else:
                c1 = ConvSpec(f'out_net.{k}.0', self.width, self.width, 0)
            c2 = ConvSpec(f'out_net.{k}.2', self.width, self.width, 1)
            self.out_specs.append((c1, c2, f'out_net.{k}.3'))
        k = self.out_blocks - 1
        if k == 0:
            self.head1 = ConvSpec(f'out_net.{k}.0', self.oc, self.width, 0, 0, self.n_streams, self.chs, self.cp)
        else:
            self.head1 = ConvSpec(f'out_net.{k}.0', self.oc, self.width, 0)
        self.head2 = ConvSpec(f'out_net.{k}.2', self.oc, self.oc, 1)

    def all_convs(self):
        seen = []
        for blocks in self.in_specs.values():
            for c1, c2, _ in blocks:
                seen += [c1, c2]
        for c1, c2, _ in self.out_specs:
            seen += [c1, c2]
        seen.append(self.head1)
        if not self.small_head:
            seen.append(self.head2)
        return seen

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    # ------------------------------------------------------------------ parameters
    def _params(self):
        # the Parameter / buffer objects of a module are stable (.to(), load_state_dict and the fused optimizer all
        # update them in place), so the name -> tensor maps are built once: walking the module tree on every call cost
        # a third of the host time of a training step
        if self._param_cache is None:
            self._param_cache = dict(self.m.named_parameters())
        return self._param_cache

    def _buffers(self):
        if self._buffer_cache is None:
            self._buffer_cache = dict(self.m.named_buffers())
        return self._buffer_cache

    def repack(self, need_dgrad):
        """(Re)build the packed bf16 operands when the canonical fp32 parameters changed."""
        params = self._params()
        version = tuple(p._version for p in params.values()) + tuple(p.data_ptr() for p in params.values()) + \
            (self.act, self.split)
        if self._pack_version is not None and self._pack_version[0] == version and \
                (self._pack_version[1] or not need_dgrad):
            return
        dev = next(iter(params.values())).device
        # every packing of the step (forward, split and data-gradient operands, padded biases) is one job of ONE
        # launch; the job table lives on the device and is rebuilt only when a buffer moves
        key = version[len(params):] + (bool(need_dgrad),)
        if key not in self._pack_jobs:
            jobs = []
            for cs in self.all_convs():
                w = params[cs.name + '.weight'].detach()
                b = params[cs.name + '.bias'].detach()
                kc = cs.kc
                if cs.w_fwd is None:
                    cs.w_fwd = torch.empty((cs.n_pad, 4 * kc * 64), dtype=torch.int16, device=dev)
                    cs.bias_pad = torch.zeros(cs.n_pad, dtype=torch.float32, device=dev)
                if self.split:
                    if cs.w_split is None:
                        cs.w_split = torch.empty((cs.n_pad, 4 * 3 * kc * 64), dtype=torch.int16, device=dev)
                        cs.unscale = torch.full((cs.n_pad,), 1.0 / SPLIT_WSCALE, dtype=torch.float32, device=dev)
                    jobs.append(PackJob(w.data_ptr(), cs.w_split.data_ptr(), 0, 0, cs.cout, cs.cin, cs.spatial, 0, cs.groups,
                                        cs.group_real, cs.group_pad, cs.n_pad, cs.cin_pad, 1, 1, SPLIT_WSCALE))
                jobs.append(PackJob(w.data_ptr(), cs.w_fwd.data_ptr(), b.data_ptr(), cs.bias_pad.data_ptr(), cs.cout, cs.cin,
                                    cs.spatial, 0, cs.groups, cs.group_real, cs.group_pad, cs.n_pad, cs.cin_pad, self.act, 0, 1.0))
                if need_dgrad:
                    kd = (cs.n_pad + 63) // 64
                    if cs.w_dgrad is None:
                        cs.w_dgrad = torch.empty((cs.cin_pad, 4 * kd * 64), dtype=torch.int16, device=dev)
                    jobs.append(PackJob(w.data_ptr(), cs.w_dgrad.data_ptr(), 0, 0, cs.cout, cs.cin, cs.spatial, 1, cs.groups,
                                        cs.group_real, cs.group_pad, cs.cin_pad, cs.n_pad, GRAD, 0, 1.0))
            arr = (PackJob * len(jobs))(*jobs)
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            max_elems = max(j.n_pad * 4 * (3 if j.split else 1) * ((j.cin_pad + 63) // 64) * 64 for j in jobs)
            self._pack_jobs[key] = (table, len(jobs), max_elems)
        table, n_jobs, max_elems = self._pack_jobs[key]
        call('mmlf_pack_conv_weights_batch', _ptr(table), n_jobs, max_elems, _stream())
        self._pack_version = (version, bool(need_dgrad))

    # ------------------------------------------------------------------ persistent scratch / gradient buffer
    def _scratch(self, geo):
        """Accumulator scratch of a step, allocated once per (device, topology): the statistics / column-sum rows the
        kernels add into (zeroed by ONE memset per pass) and the float32 per-channel rows of the BatchNorm backward."""
        key = str(self.dev)
        sc = self._scratch_cache.get(key)
        if sc is None:
            n_bn = len(self.stream_defs) * self.in_blocks + len(self.out_specs)
            n_blk = n_bn + 2
            fwd64 = torch.empty((max(n_bn, 1), 2 * 320), dtype=torch.float64, device=self.dev)
            # backward accumulators in one allocation so that one memset clears them: [3 n_blk rows of 640 doubles |
            # n_blk rows of 2 x 320 floats]
            n64, n32 = 3 * n_blk * 640, 2 * n_blk * 320
            acc = torch.empty(n64 * 8 + n32 * 4, dtype=torch.uint8, device=self.dev)
            z64 = acc[:n64 * 8].view(torch.float64).view(3 * n_blk, 640)
            z32 = acc[n64 * 8:].view(torch.float32).view(2 * n_blk, 320)
            e32 = torch.empty((n_blk, 3 * 320), dtype=torch.float32, device=self.dev)
            sc = dict(fwd64=fwd64, acc=acc, z64=z64, z32=z32, e32=e32)
            self._scratch_cache[key] = sc
        return sc

    def grad_layout(self):
        """name -> (offset, numel, shape) of every parameter in one flat float32 buffer, in named_parameters() order --
        the layout of FusedAdam's flat buffers, so a training step can write gradients straight into the optimizer's."""
        if self._glayout is None:
            off, lay = 0, {}
            for name, p in self._params().items():
                lay[name] = (off, p.numel(), tuple(p.shape))
                off += p.numel()
            self._glayout = (lay, off)
        return self._glayout

    # ------------------------------------------------------------------ kernel launch helpers
    def conv(self, geo, x, ld_in, cs, w, n_pad, cin_pad, ctype, out, ld_out, *, bias=None, scale=None, shift=None,
             relu=False, gate_bits=None, relu_bits=None, out_mode=0, n_real=0, simt=False, ab=None, out_dt=None,
             out2=None, ld_out2=0, out2_dt=None, col_sums=None, split_in=0, split_out=0, bn=None):
        a = ConvArgs()
        a.ab_dtype = self.act if ab is None else ab
        a.out_dtype = self.act if out_dt is None else out_dt
        a.out2_dtype = GRAD if out2_dt is None else out2_dt
        a.in_, a.ld_in, a.cin_pad = x.data_ptr(), ld_in, cin_pad
        a.wpack, a.n_pad = w.data_ptr(), n_pad
        a.B, a.H, a.W, a.type = geo.B, geo.H, geo.W, ctype
        a.bias = bias.data_ptr() if bias is not None else None
        a.scale = scale.data_ptr() if scale is not None else None
        a.shift = shift.data_ptr() if shift is not None else None
        a.relu = 1 if relu else 0
        a.gate_bits = gate_bits.data_ptr() if gate_bits is not None else None
        a.relu_bits = relu_bits.data_ptr() if relu_bits is not None else None
        a.ld_bits = bits_words(n_pad)
        a.out, a.ld_out, a.out_mode, a.n_real = out.data_ptr(), ld_out, out_mode, n_real
        a.out2 = out2.data_ptr() if out2 is not None else None
        a.ld_out2 = ld_out2
        a.col_sums = col_sums.data_ptr() if col_sums is not None else None
        a.split_in, a.split_out = split_in, split_out
        if bn is not None:            # BatchNorm-backward statistics fused into this (data-gradient) launch
            a.bn_z, a.ld_z, a.bn_z_dtype = bn['z'].data_ptr(), bn['ld_z'], bn['dtype']
            a.bn_scale, a.bn_shift, a.bn_mean = bn['scale'].data_ptr(), bn['shift'].data_ptr(), bn['mean'].data_ptr()
        if _lib._profile is not None:     # bench.py's per-kernel pass: which roof binds this launch
            _lib.profile_tag = 'narrow' if cin_pad <= 128 else ('wide' if n_pad > 128 else 'head')
        call('mmlf_conv2x2_simt' if simt else 'mmlf_conv2x2', C.byref(a), _stream())

    def _slots(self, geo, ch, dtype=None):
        """Activation-format slot array by default; pass GRAD or torch.float32 for the others."""
        if dtype is None:
            dtype = self.act
        if not isinstance(dtype, torch.dtype):
            dtype = _TORCH_DT[dtype]
        return torch.empty((geo.n_slots, ch), dtype=dtype, device=self.dev)

    def _bits(self, geo, n_pad):
        """ReLU sign bits of a slot array: one word per 32 channels."""
        return torch.empty((geo.n_slots, bits_words(n_pad)), dtype=torch.int32, device=self.dev)

    # ------------------------------------------------------------------ forward
    def forward(self, views, training, save, shift_disp=None):
        """views: list of (B, n, 3, H, W) fp32 CUDA tensors (h, v[, i, d]).  Returns (B, OC, H, W) fp32 and, when
        ``save`` is set, the tape needed by :meth:`backward`.  ``shift_disp`` fuses the ESE Shift into the packing.

        With ``save`` every activation that the backward pass multiplies on the tensor cores (block inputs and the
        ReLU'd output of each first conv) is additionally written in the gradient format (bf16) by the kernel that
        produces it -- tcgen05 kind::f16 needs both operands of the weight-gradient GEMM in one format -- and the
        fp16 copies are dropped as soon as the next layer has consumed them."""
        _lib.require_device()
        # validate BEFORE the first launch: a kernel handed host pointers faults asynchronously and poisons the context
        for v in views:
            if not (isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
                raise RuntimeError('mmlf_b200: view stacks must be contiguous float32 CUDA tensors (feed_forward.py:226-232 '
                                   'uses .view); there is no CPU path')
        p0 = next(iter(self._params().values()))
        if p0.device != views[0].device:
            raise RuntimeError(f'mmlf_b200: the model lives on {p0.device} but the inputs on {views[0].device}; '
                               'move the module with .cuda() / .to(device) first (there is no CPU path)')
        if self.split:
            if save or training:
                raise RuntimeError("precision='split' is an inference mode (model.eval(), torch.no_grad())")
            return self._forward_split(views, shift_disp), None
        h = views[0]
        B, n, c3, H, W = h.shape
        self.dev = h.device
        geo = Geometry(B, H, W)
        st = _stream()
        self.repack(need_dgrad=save)
        bn_train = training and self.has_bn
        tape = {'geo': geo, 'streams': {}, 'out': [], 'bn_train': bn_train,
                'act_dt': self.act if self.mixed_wgrad else GRAD} if save else None
        bufs = self._buffers()
        params = self._params()
        cin0_pad = pad16(n * c3)
        dual = save and self.act != GRAD and not self.mixed_wgrad   # bf16 twins of the backward pass's MMA operands

        # per-forward scratch for the BatchNorm layers, allocated once (one launch each instead of five per block)
        n_bn = len(self.stream_defs) * self.in_blocks + len(self.out_specs)
        self._fwd_sums = None
        if bn_train:
            self._fwd_sums = self._scratch(geo)['fwd64']
            _zero(self._fwd_sums)
        self._fwd_consts = torch.empty((n_bn, 4, 320), dtype=torch.float32, device=self.dev) if bn_train else None
        self._fwd_bn_idx = 0
        feats = self._slots(geo, self.feat_ld)
        featsg = self._slots(geo, self.feat_ld, GRAD) if dual else (feats if save else None)
        # all view stacks -> slot layout in ONE launch (grid.z = stack), the bf16 twins written from the same read;
        # with `shift_disp` the ESE Shift is fused into it
        ns = len(self.stream_defs)
        xs = [self._slots(geo, cin0_pad) for _ in range(ns)]
        xgs = [self._slots(geo, cin0_pad, GRAD) for _ in range(ns)] if dual else (xs if save else [None] * ns)
        if B * (H + 1) <= 65535:
            PtrArr, IntArr = C.c_void_p * ns, C.c_int * ns
            call('mmlf_pack_stacks', PtrArr(*[views[si].data_ptr() for si in range(ns)]), IntArr(*range(ns)), ns, B, n, H,
                 W, PtrArr(*[t.data_ptr() for t in xs]), PtrArr(*[t.data_ptr() for t in xgs]) if dual else None, cin0_pad,
                 self.act, GRAD, 0 if shift_disp is None else 1, 0.0 if shift_disp is None else float(shift_disp), st)
        else:                                              # very large batches: one (slab-splitting) launch per stack
            for si in range(ns):
                for dst, dt in ((xs[si], self.act),) + (((xgs[si], GRAD),) if dual else ()):
                    if shift_disp is None:
                        call('mmlf_pack_views', _ptr(views[si]), B, n * c3, H, W, _ptr(dst), cin0_pad, dt, st)
                    else:
                        call('mmlf_shift_pack', _ptr(views[si]), si, B, n, H, W, float(shift_disp), _ptr(dst), cin0_pad, dt, st)
        for si, (key, net, spatial) in enumerate(self.stream_defs):
            x, xg = xs[si], xgs[si]
            ld_x = cin0_pad
            recs = []
            blocks = self.in_specs[key]
            for k, (c1, c2, bnp) in enumerate(blocks):
                last = k == len(blocks) - 1
                if last:      # write straight into this stream's slice of the concatenated feature buffer
                    y, ld_y = feats[:, si * self.cp:], self.feat_ld
                    yg = featsg[:, si * self.cp:] if save else None
                else:
                    y, ld_y = self._slots(geo, c2.n_pad), c2.n_pad
                    yg = self._slots(geo, c2.n_pad, GRAD) if dual else (y if save else None)
                rec = self._block_fwd(geo, x, xg, ld_x, c1, c2, bnp, y, yg, ld_y, training, bn_train, save, bufs, params)
                recs.append(rec)
                x, xg, ld_x = y, yg, ld_y
            if save:
                tape['streams'][key] = recs
        x, xg, ld_x = feats, featsg, self.feat_ld
        for k, (c1, c2, bnp) in enumerate(self.out_specs):
            y = self._slots(geo, c2.n_pad)
            yg = self._slots(geo, c2.n_pad, GRAD) if dual else (y if save else None)
            rec = self._block_fwd(geo, x, xg, ld_x, c1, c2, bnp, y, yg, c2.n_pad, training, bn_train, save, bufs, params)
            if save:
                tape['out'].append(rec)
            x, xg, ld_x = y, yg, c2.n_pad
        # head block: conv -> ReLU -> conv, no BN / ReLU after (feed_forward.py:185)
        h1 = self.head1
        out = torch.empty((B, self.oc, H, W), dtype=torch.float32, device=self.dev)
        midg = bits = None
        if self.small_head:
            mid = self._slots(geo, h1.n_pad, torch.float32)
            self.conv(geo, x, ld_x, h1, h1.w_fwd, h1.n_pad, h1.cin_pad, 0, mid, h1.n_pad, bias=h1.bias_pad, relu=True,
                      out_mode=1)
            w2 = params[self.head2.name + '.weight'].detach()
            b2 = params[self.head2.name + '.bias'].detach()
            call('mmlf_head_small', _ptr(mid), h1.n_pad, self.oc, _ptr(w2), _ptr(b2), B, H, W, _ptr(out), st)
        else:
            h2 = self.head2
            mid = self._slots(geo, h1.n_pad)
            if save:
                midg = self._slots(geo, h1.n_pad, GRAD) if dual else mid
                bits = self._bits(geo, h1.n_pad)
            self.conv(geo, x, ld_x, h1, h1.w_fwd, h1.n_pad, h1.cin_pad, 0, mid, h1.n_pad, bias=h1.bias_pad, relu=True,
                      out2=midg if dual else None, ld_out2=h1.n_pad, relu_bits=bits)
            self.conv(geo, mid, h1.n_pad, h2, h2.w_fwd, h2.n_pad, h2.cin_pad, 1, out, 0, bias=h2.bias_pad,
                      out_mode=2, n_real=self.oc)
        if save:
            tape['head'] = {'xg': xg, 'ld_x': ld_x, 'mid': mid if self.small_head else None, 'midg': midg, 'bits': bits}
        return out, tape

    def _forward_split(self, views, shift_disp=None):
        """Eval-mode forward in split precision.  Slot arrays hold [hi | lo] blocks: an array of C channels has pitch
        2 * C with hi in columns [0, C) and lo in [C, 2 C)."""
        if shift_disp is not None:                      # ESE member: resample in fp32 first (bit-exact Shift kernel)
            from . import ops
            full = list(views) + [views[0]] * (4 - len(views))
            views = ops.lf_shift(*full, float(shift_disp))[:len(views)]
        h = views[0]
        B, n, c3, H, W = h.shape
        self.dev = h.device
        geo = Geometry(B, H, W)
        st = _stream()
        self.repack(need_dgrad=False)
        bufs, params = self._buffers(), self._params()
        cin0_pad = pad16(n * c3)
        half = torch.float16

        def slots2(ch):
            return torch.empty((geo.n_slots, 2 * ch), dtype=half, device=self.dev)

        def block(x, ld_x, lo_x, c1, c2, bnp, y, ld_y, lo_y):
            a1 = slots2(c1.n_pad)
            self.conv(geo, x, ld_x, c1, c1.w_split, c1.n_pad, c1.cin_pad, 0, a1, 2 * c1.n_pad, scale=c1.unscale,
                      shift=c1.bias_pad, relu=True, ab=FP16, out_dt=FP16, split_in=lo_x, split_out=c1.n_pad)
            if self.has_bn:
                gamma, beta = params[bnp + '.weight'].detach(), params[bnp + '.bias'].detach()
                scale, shift = self._eval_fold(bnp, c2, gamma, beta, bufs[bnp + '.running_mean'], bufs[bnp + '.running_var'])
                self.conv(geo, a1, 2 * c1.n_pad, c2, c2.w_split, c2.n_pad, c2.cin_pad, 1, y, ld_y,
                          scale=scale * (1.0 / SPLIT_WSCALE), shift=shift, relu=True, ab=FP16, out_dt=FP16,
                          split_in=c1.n_pad, split_out=lo_y)
            else:
                self.conv(geo, a1, 2 * c1.n_pad, c2, c2.w_split, c2.n_pad, c2.cin_pad, 1, y, ld_y, scale=c2.unscale,
                          shift=c2.bias_pad, relu=True, ab=FP16, out_dt=FP16, split_in=c1.n_pad, split_out=lo_y)

        feats = slots2(self.feat_ld)                    # hi blocks of the streams in [0, feat_ld), lo blocks behind
        for si, (key, net, spatial) in enumerate(self.stream_defs):
            v = views[si]
            assert v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()
            x = slots2(cin0_pad)
            call('mmlf_pack_views_split', _ptr(v), B, n * c3, H, W, _ptr(x), 2 * cin0_pad, cin0_pad, st)
            ld_x, lo_x = 2 * cin0_pad, cin0_pad
            blocks = self.in_specs[key]
            for k, (c1, c2, bnp) in enumerate(blocks):
                if k == len(blocks) - 1:
                    y, ld_y, lo_y = feats[:, si * self.cp:], 2 * self.feat_ld, self.feat_ld
                else:
                    y, ld_y, lo_y = slots2(c2.n_pad), 2 * c2.n_pad, c2.n_pad
                block(x, ld_x, lo_x, c1, c2, bnp, y, ld_y, lo_y)
                x, ld_x, lo_x = y, ld_y, lo_y
        x, ld_x, lo_x = feats, 2 * self.feat_ld, self.feat_ld
        for c1, c2, bnp in self.out_specs:
            y = slots2(c2.n_pad)
            block(x, ld_x, lo_x, c1, c2, bnp, y, 2 * c2.n_pad, c2.n_pad)
            x, ld_x, lo_x = y, 2 * c2.n_pad, c2.n_pad
        h1 = self.head1
        out = torch.empty((B, self.oc, H, W), dtype=torch.float32, device=self.dev)
        if self.small_head:
            mid = self._slots(geo, h1.n_pad, torch.float32)
            self.conv(geo, x, ld_x, h1, h1.w_split, h1.n_pad, h1.cin_pad, 0, mid, h1.n_pad, scale=h1.unscale,
                      shift=h1.bias_pad, relu=True, out_mode=1, ab=FP16, split_in=lo_x)
            w2 = params[self.head2.name + '.weight'].detach()
            b2 = params[self.head2.name + '.bias'].detach()
            call('mmlf_head_small', _ptr(mid), h1.n_pad, self.oc, _ptr(w2), _ptr(b2), B, H, W, _ptr(out), st)
        else:
            h2 = self.head2
            mid = slots2(h1.n_pad)
            self.conv(geo, x, ld_x, h1, h1.w_split, h1.n_pad, h1.cin_pad, 0, mid, 2 * h1.n_pad, scale=h1.unscale,
                      shift=h1.bias_pad, relu=True, ab=FP16, out_dt=FP16, split_in=lo_x, split_out=h1.n_pad)
            self.conv(geo, mid, 2 * h1.n_pad, h2, h2.w_split, h2.n_pad, h2.cin_pad, 1, out, 0, scale=h2.unscale,
                      shift=h2.bias_pad, out_mode=2, n_real=self.oc, ab=FP16, split_in=h1.n_pad)
        return out

    def _block_fwd(self, geo, x, xg, ld_x, c1, c2, bnp, y, yg, ld_y, training, bn_train, save, bufs, params):
        """conv(k2,p1) -> ReLU -> conv(k2,p0) [-> BN] -> ReLU   (feed_forward.py:122-137).  x / y are the block input
        and output in the activation format, xg / yg their gradient-format twins (None unless ``save``)."""
        st = _stream()
        dual = save and self.act != GRAD and not self.mixed_wgrad
        a1 = self._slots(geo, c1.n_pad)
        a1g = bits = None
        if save:
            a1g = self._slots(geo, c1.n_pad, GRAD) if dual else a1
            bits = self._bits(geo, c1.n_pad)
        self.conv(geo, x, ld_x, c1, c1.w_fwd, c1.n_pad, c1.cin_pad, 0, a1, c1.n_pad, bias=c1.bias_pad, relu=True,
                  out2=a1g if dual else None, ld_out2=c1.n_pad, relu_bits=bits)
        rec = {'xg': xg, 'ld_x': ld_x, 'a1g': a1g, 'bits': bits, 'yg': yg, 'ld_y': ld_y, 'c1': c1, 'c2': c2,
               'bnp': bnp} if save else None
        C_real, Cp = c2.cout, c2.n_pad
        if not self.has_bn:
            self.conv(geo, a1, c1.n_pad, c2, c2.w_fwd, Cp, c2.cin_pad, 1, y, ld_y, bias=c2.bias_pad, relu=True,
                      out2=yg if dual else None, ld_out2=ld_y)
            return rec
        gamma, beta = params[bnp + '.weight'].detach(), params[bnp + '.bias'].detach()
        rmean, rvar = bufs[bnp + '.running_mean'], bufs[bnp + '.running_var']
        if not bn_train and not save:
            # eval: BN folded into the conv epilogue, y = relu(acc * scale + shift)
            scale, shift = self._eval_fold(bnp, c2, gamma, beta, rmean, rvar)
            self.conv(geo, a1, c1.n_pad, c2, c2.w_fwd, Cp, c2.cin_pad, 1, y, ld_y, scale=scale, shift=shift, relu=True)
            return rec
        if not bn_train:
            # --train_eval_mode (train/cli.py:227-230): gradients through a BatchNorm that normalises with its RUNNING
            # statistics.  Same tape as the training path (z kept, ReLU mask recomputed from it), the constants come
            # from the running statistics and the backward pass drops the batch-mean terms (bn_bwd_apply train = 0).
            scale, shift, save_mean, save_invstd = self._eval_consts(bnp, c2, gamma, beta, rmean, rvar)
            z = self._slots(geo, Cp)
            self.conv(geo, a1, c1.n_pad, c2, c2.w_fwd, Cp, c2.cin_pad, 1, z, Cp, bias=c2.bias_pad)
            call('mmlf_bn_apply_relu', _ptr(z), Cp, _ptr(scale), _ptr(shift), Cp, geo.B, geo.H, geo.W, self.act, _ptr(y),
                 ld_y, _ptr(yg) if dual else C.c_void_p(0), ld_y, GRAD, st)
            rec.update(z=z, scale=scale, shift=shift, save_mean=save_mean, save_invstd=save_invstd, bn_eval=True)
            return rec
        k = self._fwd_bn_idx
        self._fwd_bn_idx += 1
        consts = self._fwd_consts[k]
        scale, shift, save_mean, save_invstd = consts[0, :Cp], consts[1, :Cp], consts[2, :Cp], consts[3, :Cp]
        z = self._slots(geo, Cp)
        # batch statistics (sum, sum of squares of the stored z) come out of the conv epilogue
        sums = self._fwd_sums[k, :2 * Cp]
        self.conv(geo, a1, c1.n_pad, c2, c2.w_fwd, Cp, c2.cin_pad, 1, z, Cp, bias=c2.bias_pad, col_sums=sums)
        nbt = bufs.get(bnp + '.num_batches_tracked')
        call('mmlf_bn_finalize', _ptr(sums), C_real, Cp, geo.count, _ptr(gamma), _ptr(beta), _ptr(rmean), _ptr(rvar),
             _ptr(nbt), float(self.m.batchnorm_momentum), float(self.m.bn_eps), _ptr(scale), _ptr(shift),
             _ptr(save_mean), _ptr(save_invstd), st)
        torch.autograd.graph.increment_version([t for t in (rmean, rvar, nbt) if t is not None])   # raw-pointer writes
        call('mmlf_bn_apply_relu', _ptr(z), Cp, _ptr(scale), _ptr(shift), Cp, geo.B, geo.H, geo.W, self.act, _ptr(y),
             ld_y, _ptr(yg) if dual else C.c_void_p(0), ld_y, GRAD, st)
        if save:
            rec.update(z=z, scale=scale, shift=shift, save_mean=save_mean, save_invstd=save_invstd)
        return rec

    def _eval_fold(self, bnp, c2, gamma, beta, rmean, rvar):
        """Eval-mode BN folded to per-channel scale / shift; cached until a parameter or running statistic changes."""
        cbias = self._params()[c2.name + '.bias']
        key = (gamma._version, beta._version, rmean._version, rvar._version, cbias._version, gamma.data_ptr(),
               rmean.data_ptr(), cbias.data_ptr())
        hit = self._fold_cache.get(bnp)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        Cp = c2.n_pad
        scale = torch.empty(Cp, dtype=torch.float32, device=self.dev)
        shift = torch.empty(Cp, dtype=torch.float32, device=self.dev)
        call('mmlf_bn_fold_eval', c2.cout, Cp, _ptr(gamma), _ptr(beta), _ptr(rmean), _ptr(rvar),
             _ptr(c2.bias_pad), float(self.m.bn_eps), _ptr(scale), _ptr(shift), _stream())
        self._fold_cache[bnp] = (key, scale, shift)
        return scale, shift

    def _eval_consts(self, bnp, c2, gamma, beta, rmean, rvar):
        """Eval-mode BN constants for a differentiable forward: scale / shift applied to z (conv bias included in z),
        and the running mean / inverse standard deviation on the padded channel pitch."""
        Cp, Cr = c2.n_pad, c2.cout
        consts = torch.zeros((4, Cp), dtype=torch.float32, device=self.dev)
        scale, shift, mean, invstd = consts[0], consts[1], consts[2], consts[3]
        call('mmlf_bn_fold_eval', Cr, Cp, _ptr(gamma), _ptr(beta), _ptr(rmean), _ptr(rvar), C.c_void_p(0),
             float(self.m.bn_eps), _ptr(scale), _ptr(shift), _stream())
        mean[:Cr].copy_(rmean)
        invstd[:Cr].copy_(torch.rsqrt(rvar + float(self.m.bn_eps)))
        return scale, shift, mean, invstd

    # ------------------------------------------------------------------ backward
    def backward(self, tape, g_out, flat=None):
        """g_out: (B, OC, H, W) fp32.  Every parameter gradient is written into one flat float32 buffer laid out like
        :meth:`grad_layout` -- ``flat`` if given (the optimizer's gradient buffer: no copies, nothing returned through
        autograd), else a buffer owned by the engine.  Returns {param name: view into that buffer}.

        No framework kernel runs here: accumulator scratch is cleared by one memset, weight gradients are reduced
        straight into their slice, BatchNorm gamma / beta gradients are written (or accumulated, for the shared in-nets)
        by the kernel that computes them, and all the short bias-gradient vectors are finished by ONE launch over a
        device job table at the end (``mmlf_vec_jobs``)."""
        geo = tape['geo']
        st = _stream()
        params = self._params()
        dev = self.dev
        layout, total = self.grad_layout()
        if flat is None:
            # the previous pass's buffer may still be some parameter's .grad (autograd adopted our views): never
            # overwrite it, hand out a new one instead
            lo = self._gflat.data_ptr() if self._gflat is not None else 0
            hi = lo + total * 4
            if self._gflat is None or self._gflat.device != dev or \
                    any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params.values()):
                self._gflat = torch.empty(total, dtype=torch.float32, device=dev)
            flat = self._gflat
            _zero(flat)
            written = None
        else:
            assert flat.numel() == total and flat.dtype == torch.float32 and flat.is_contiguous()
            written = set()          # the caller's buffer holds earlier gradients: accumulate from the first write on
        base = flat.data_ptr()

        def gptr(name):
            return C.c_void_p(base + 4 * layout[name][0])

        seen = set()

        def first(name):
            """True when this is the first write to `name` in this pass AND the buffer was cleared for it."""
            new = name not in seen
            seen.add(name)
            return new and written is None

        if self._ws is None or self._ws.device != dev:
            ws_bytes = max(_lib.lib().mmlf_conv2x2_wgrad_workspace(cs.n_pad, cs.cin_pad) for cs in self.all_convs())
            self._ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
        ws = self._ws
        sc = self._scratch(geo)
        z64, z32, e32 = sc['z64'], sc['z32'], sc['e32']
        _zero(sc['acc'])
        pool = {'z64': 0, 'z32': 0, 'e32': 0}
        jobs = {}                                       # dst name -> [n, accumulate into dst, (src ptr, is f64), ...]

        def take(name, buf, n):
            i = pool[name]
            pool[name] = i + 1
            return buf[i, :n]

        # Weight gradients are off the critical path of the backward chain (nothing downstream reads them), so they CAN
        # run on a side stream, overlapping the tensor-bound wgrad kernels of block k with the HBM-bound BatchNorm
        # passes of block k - 1.  Opt-in (MMLF_OVERLAP_WGRAD=1): measured on B200 it gains 2 % at 64 patches per GPU and
        # loses 2 % at 512 -- the step runs into the board power cap (SM clock ~1.5 GHz), so concurrency only trades
        # clock for occupancy.
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap_wgrad and not torch.cuda.is_current_stream_capturing() else None
        held, pending = [], []                         # tensors the side stream still reads: [(event, tensors)] per block

        def end_block():
            """Bound the lag of the side stream to one block: its tensors are kept alive until the main stream has
            waited for the side stream's work of the PREVIOUS block (cheaper and tighter on memory than
            Tensor.record_stream, which at B = 512 drove the caching allocator into cudaMalloc retries)."""
            if side is None:
                return
            ev = torch.cuda.Event()
            ev.record(side)
            pending.append((ev, list(held)))
            held.clear()
            if len(pending) > 1:
                ev0, hold0 = pending.pop(0)
                main.wait_event(ev0)
                hold0.clear()

        def conv_param_grads(cs, dout, ld_dout, actg, ld_act, dbias=None):
            """dW via the tcgen05 wgrad kernel (both operands in the gradient format), reduced straight into the
            parameter's slice of the flat buffer; db from ``dbias`` = (scratch row, is_float64) when the kernel that
            produced ``dout`` already summed its columns, else via a column-sum pass into a scratch row.  Accumulates
            for modules that are called twice per forward."""
            wname, bname = cs.name + '.weight', cs.name + '.bias'
            acc = 0 if first(wname) else 1
            first(bname)
            if dbias is None:
                row = take('z32', z32, cs.n_pad)
                dbias = (row, False)

                def colsum(sst):
                    call('mmlf_colsum16', _ptr(dout), ld_dout, cs.n_pad, geo.n_slots, GRAD, _ptr(row), 1, sst)
            else:
                colsum = None
            # one job per destination (jobs of a launch run concurrently): a second call of a shared module adds a source
            jobs.setdefault(bname, [cs.cout, 0 if written is None else 1]).append((dbias[0].data_ptr(), bool(dbias[1])))

            def work():
                sst = _stream()
                if _lib._profile is not None:
                    _lib.profile_tag = 'narrow' if cs.cin_pad <= 128 else ('wide' if cs.n_pad > 128 else 'head')
                call('mmlf_conv2x2_wgrad_canonical', _ptr(dout), ld_dout, cs.n_pad, _ptr(actg), ld_act, cs.cin_pad, geo.B,
                     geo.H, geo.W, cs.type, tape['act_dt'], GRAD, _ptr(ws), cs.cout, cs.cin, cs.spatial, cs.groups, cs.group_real,
                     cs.group_pad, gptr(wname), acc, sst)
                if colsum is not None:
                    colsum(sst)
            if side is None:
                work()
                return
            side.wait_stream(main)                         # producers of dout / actg / dbias, allocation of the grads
            with torch.cuda.stream(side):
                work()
            held.extend(t for t in (dout, actg) if t is not None)

        def dgrad(cs, dout, ld_dout, out, ld_out, gate_bits=None, col_sums=None, bn=None):
            """Data gradient: the conv kernel of the other type with rotated, transposed weights."""
            self.conv(geo, dout, ld_dout, cs, cs.w_dgrad, cs.cin_pad, cs.n_pad, 1 - cs.type, out, ld_out,
                      gate_bits=gate_bits, ab=GRAD, out_dt=GRAD, col_sums=col_sums, bn=bn)

        def bn_of(prev_rec, cin_pad):
            """BatchNorm whose output feeds a conv with `cin_pad` input channels: its backward statistics can come out
            of that conv's data-gradient epilogue (saves one read of the gradient and of z per BN layer)."""
            if not self.fuse_bn_bwd or prev_rec is None or 'z' not in prev_rec or prev_rec['c2'].n_pad != cin_pad \
                    or prev_rec.get('bn_eval'):
                return None, None
            if cin_pad < 256:
                # narrow layers are epilogue bound already: measured 0.075 -> 0.105 ms for the 70-channel data gradient
                # against 0.048 ms for the standalone reduction (wide layers: 0.293 -> 0.382 ms against 0.141 ms)
                return None, None
            psums = take('z64', z64, 2 * cin_pad)
            return psums, dict(z=prev_rec['z'], ld_z=prev_rec['c2'].n_pad, dtype=self.act, scale=prev_rec['scale'],
                               shift=prev_rec['shift'], mean=prev_rec['save_mean'])

        # ---- head
        hd = tape['head']
        h1 = self.head1
        g_x = self._slots(geo, h1.cin_pad, GRAD)
        if self.small_head:
            h2n = self.head2.name
            w2 = params[h2n + '.weight'].detach()
            gmid = self._slots(geo, h1.n_pad, GRAD)
            first(h2n + '.weight'), first(h2n + '.bias')      # head_small_bwd adds into its (cleared) outputs
            call('mmlf_head_small_bwd', _ptr(g_out), _ptr(hd['mid']), h1.n_pad, self.oc, _ptr(w2), geo.B, geo.H, geo.W,
                 _ptr(gmid), h1.n_pad, gptr(h2n + '.weight'), gptr(h2n + '.bias'), st)
            conv_param_grads(h1, gmid, h1.n_pad, hd['xg'], hd['ld_x'])
        else:
            h2 = self.head2
            gz = self._slots(geo, h2.n_pad, GRAD)
            call('mmlf_pack_views', _ptr(g_out), geo.B, self.oc, geo.H, geo.W, _ptr(gz), h2.n_pad, GRAD, st)
            conv_param_grads(h2, gz, h2.n_pad, hd['midg'], h1.n_pad)
            gmid = self._slots(geo, h1.n_pad, GRAD)
            sums = take('z64', z64, 2 * h1.n_pad)
            dgrad(h2, gz, h2.n_pad, gmid, h1.n_pad, gate_bits=hd['bits'], col_sums=sums)
            conv_param_grads(h1, gmid, h1.n_pad, hd['xg'], hd['ld_x'], dbias=(sums, True))
        pre, bn = bn_of(tape['out'][-1] if tape['out'] else None, h1.cin_pad)
        dgrad(h1, gmid, h1.n_pad, g_x, h1.cin_pad, col_sums=pre, bn=bn)
        gy, ld_gy = g_x, h1.cin_pad
        end_block()

        def block_bwd(rec, gy, ld_gy, need_gx, pre_sums=None, prev_rec=None):
            """Backward of one block.  pre_sums: BatchNorm-backward statistics of this block that the producer of gy
            already accumulated; prev_rec: the block whose output this block reads.  Returns (gx, statistics for the
            previous block or None)."""
            c1, c2, bnp = rec['c1'], rec['c2'], rec['bnp']
            Cp, C_real = c2.n_pad, c2.cout
            dz = self._slots(geo, Cp, GRAD)
            db2 = None
            if self.has_bn:
                if pre_sums is None:
                    sums = take('z64', z64, 2 * Cp)
                    call('mmlf_bn_bwd_reduce', _ptr(gy), ld_gy, _ptr(rec['z']), Cp, _ptr(rec['scale']),
                         _ptr(rec['shift']), _ptr(rec['save_mean']), _ptr(rec['save_invstd']), Cp, geo.B, geo.H, geo.W,
                         GRAD, self.act, _ptr(sums), st)
                else:
                    sums = pre_sums
                gamma = params[bnp + '.weight'].detach()
                acc = 0 if first(bnp + '.weight') else 1
                first(bnp + '.bias')
                fsums = take('e32', e32, 3 * Cp)
                db2 = take('z32', z32, Cp)
                call('mmlf_bn_bwd_apply', _ptr(gy), ld_gy, _ptr(rec['z']), Cp, _ptr(rec['scale']), _ptr(rec['shift']),
                     _ptr(gamma), _ptr(rec['save_mean']), _ptr(rec['save_invstd']), _ptr(sums), geo.count,
                     0 if rec.get('bn_eval') else (1 if pre_sums is None else 2), C_real, Cp,
                     geo.B, geo.H, geo.W, GRAD, self.act, _ptr(dz), Cp, gptr(bnp + '.weight'), gptr(bnp + '.bias'), acc,
                     _ptr(fsums), _ptr(db2), st)
                db2 = (db2, False)
            else:
                call('mmlf_relu_bwd', _ptr(gy), ld_gy, _ptr(rec['yg']), rec['ld_y'], Cp, geo.n_slots, GRAD, tape['act_dt'],
                     _ptr(dz), Cp, st)
            # data gradients first (critical path), then the weight gradients of the same operands on the side stream
            da1 = self._slots(geo, c1.n_pad, GRAD)
            sums1 = take('z64', z64, 2 * c1.n_pad)
            dgrad(c2, dz, Cp, da1, c1.n_pad, gate_bits=rec['bits'], col_sums=sums1)
            conv_param_grads(c2, dz, Cp, rec['a1g'], c1.n_pad, dbias=db2)
            gx = psums = None
            if need_gx:
                gx = self._slots(geo, c1.cin_pad, GRAD)
                psums, bn = bn_of(prev_rec, c1.cin_pad)
                dgrad(c1, da1, c1.n_pad, gx, c1.cin_pad, col_sums=psums, bn=bn)
            conv_param_grads(c1, da1, c1.n_pad, rec['xg'], rec['ld_x'], dbias=(sums1, True))
            end_block()
            return gx, psums

        outs = tape['out']
        for k in reversed(range(len(outs))):
            gy, pre = block_bwd(outs[k], gy, ld_gy, True, pre_sums=pre, prev_rec=outs[k - 1] if k > 0 else None)
            ld_gy = outs[k]['c1'].cin_pad
        # gy is now the gradient of the concatenated feature buffer [n_slots][feat_ld]
        for si, (key, net, spatial) in enumerate(self.stream_defs):
            g, ld_g = gy[:, si * self.cp:], ld_gy
            recs = tape['streams'][key]
            pre = None                          # the feature-buffer gradient covers four BN layers: standalone reduce
            for j in reversed(range(len(recs))):
                g, pre = block_bwd(recs[j], g, ld_g, need_gx=(j != 0), pre_sums=pre, prev_rec=recs[j - 1] if j > 0 else None)
                ld_g = recs[j]['c1'].cin_pad
        if side is not None:
            main.wait_stream(side)
        # ---- all the short vectors in one launch; the device job table is rebuilt only when a pointer moved
        key = []
        for name, (n, acc, *srcs) in jobs.items():
            assert len(srcs) <= 2, 'a module is called at most twice per forward'
            s2 = srcs[1] if len(srcs) > 1 else (0, False)
            key.append((srcs[0][0], s2[0], base + 4 * layout[name][0], n, int(srcs[0][1]) | (int(s2[1]) << 1), acc))
        key = tuple(key)
        if key not in self._vec_jobs:
            if len(self._vec_jobs) >= 8 and not torch.cuda.is_current_stream_capturing():
                # eager passes that keep moving their gradient buffer: drop the oldest tables, but never one that a live
                # captured step may address (those are registered in _pinned_tables)
                for old in [k for k in self._vec_jobs if k not in self._pinned_tables][:4]:
                    del self._vec_jobs[old]
            arr = (VecJob * len(key))(*[VecJob(s1, s2 or None, dst, n, f64, acc, 0) for s1, s2, dst, n, f64, acc in key])
            self._vec_jobs[key] = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        if torch.cuda.is_current_stream_capturing():
            self._pinned_tables.add(key)
        call('mmlf_vec_jobs', _ptr(self._vec_jobs[key]), len(key), st)
        return {name: flat[off:off + n].view(shape) for name, (off, n, shape) in layout.items()}
