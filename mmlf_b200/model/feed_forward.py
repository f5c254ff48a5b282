"""Drop-in for ``mmlf.model.feed_forward`` (/root/reference/mmlf/model/feed_forward.py).

Same constructor keywords, same ``state_dict`` layout (the parameter containers are the
same ``nn.Sequential`` of ``nn.Conv2d`` / ``nn.ReLU`` / ``nn.BatchNorm2d`` modules, created in the same order so
that a given ``torch.manual_seed`` yields the same initial weights), same output dict.  The modules are parameter
containers only: ``forward`` runs the hand-written sm_100a kernels through :class:`mmlf_b200.engine.Engine`.
"""
import torch
import torch.nn as nn

from .. import ops
from ..engine import Engine


def laplacian(x, mu, b):
    """feed_forward.py:9-12 (kept for API parity; the fused kernel is ``ops.upr_posterior``)."""
    mu = mu.unsqueeze(1)
    b = b.unsqueeze(1)
    return 1.0 / (2.0 * b) * torch.exp(-torch.abs(x - mu) / b)


class _NetFunction(torch.autograd.Function):
    """Whole-network forward/backward as one autograd node: the backward pass is the hand-written chain in
    Engine.backward, not autograd over per-layer ops."""

    @staticmethod
    def forward(ctx, engine, training, save, n_views, *tensors):
        views = tensors[:n_views]
        out, tape = engine.forward(list(views), training, save=save)
        ctx.engine, ctx.tape, ctx.n_views = engine, tape, n_views
        ctx.names = [n for n, _ in engine.m.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.tape is None:
            raise RuntimeError('backward through a forward pass that did not record a tape')
        grads = ctx.engine.backward(ctx.tape, g_out.contiguous())
        ctx.tape = None
        return (None, None, None, None) + (None,) * ctx.n_views + tuple(grads.get(n) for n in ctx.names)


class LazyOutputs(dict):
    """Output dict whose heavy side outputs (UPR posterior, DPP one_hot/posterior: SURVEY.md H7) are computed on
    first access.  Behaves like the reference's plain dict for every read access."""

    def __init__(self, eager, lazy):
        super().__init__(eager)
        self._lazy = dict(lazy)
        for k in lazy:
            super().__setitem__(k, None)

    def _resolve(self, k):
        if k in self._lazy:
            for kk, vv in self._lazy.pop(k)().items():
                self._lazy.pop(kk, None)
                super().__setitem__(kk, vv)

    def __getitem__(self, k):
        self._resolve(k)
        return super().__getitem__(k)

    def get(self, k, default=None):
        if k not in self:
            return default
        return self[k]

    def items(self):
        for k in list(self._lazy):
            self._resolve(k)
        return super().items()

    def values(self):
        for k in list(self._lazy):
            self._resolve(k)
        return super().values()


class FeedForward(nn.Module):
    """EPINET-style multi-stream FCN (feed_forward.py:15) on B200 kernels."""

    def __init__(self, model_ksize, model_in_blocks, model_out_blocks, model_chs, model_views, model_cross,
                 model_uncert, model_unet, model_discrete, model_no_batchnorm, model_batchnorm_momentum,
                 val_disp_min, val_disp_max, **kwargs):
        super(FeedForward, self).__init__()
        if model_ksize < 1 or model_ksize > 7:
            raise NotImplementedError('model_ksize must be in 1..7: 2 is the published topology (tensor-core path), the '
                                      'others run on the float32 layer kernels (SURVEY.md section 8f.4)')
        if model_unet and model_discrete:
            raise NotImplementedError('--model_unet has a 1- or 2-channel head (feed_forward.py:196-202)')
        self.ksize = model_ksize
        self.unet = bool(model_unet)
        self.chs = model_chs
        self.views = model_views
        self.cross = model_cross
        self.uncert = model_uncert
        self.discrete = model_discrete
        self.no_batchnorm = model_no_batchnorm
        self.batchnorm_momentum = model_batchnorm_momentum
        self.bn_eps = 1e-5
        self.disp_min = val_disp_min
        self.disp_max = val_disp_max
        self.steps = 4
        if model_cross:
            self.steps = 2
        self.steps *= model_views * 3                                   # feed_forward.py:81-84
        self.padding1 = model_ksize // 2                                 # feed_forward.py:86-92
        self.padding2 = model_ksize // 2 if model_ksize % 2 == 1 else model_ksize // 2 - 1
        self.n_in_blocks = model_in_blocks
        self.n_out_blocks = model_out_blocks

        self.in_net_hv = self.init_in_net(model_in_blocks)
        if not model_cross:
            self.in_net_id = self.init_in_net(model_in_blocks)
        self.out_net = self.init_unet() if model_unet else self.init_out_net(model_out_blocks)   # feed_forward.py:99-102
        self._engine = None
        self._bins = {}
        self._graphs = {}
        self.use_cuda_graph = True        # inference launches are replayed from a captured CUDA graph

    # -- parameter containers, same construction order as the reference (feed_forward.py:104-187)
    def block(self, ch_in, ch_out=None, out_bn_relu=True):
        if ch_out is None:
            ch_out = ch_in
        layers = [nn.Conv2d(ch_in, ch_out, self.ksize, padding=self.padding1), nn.ReLU(),
                  nn.Conv2d(ch_out, ch_out, self.ksize, padding=self.padding2)]
        if out_bn_relu:
            if not self.no_batchnorm:
                layers.append(nn.BatchNorm2d(ch_out, momentum=self.batchnorm_momentum))
            layers.append(nn.ReLU())
        return nn.Sequential(*layers)

    def init_in_net(self, n_blocks):
        assert n_blocks >= 1
        blocks = [self.block(self.views * 3, self.chs)]
        for _ in range(n_blocks - 1):
            blocks.append(self.block(self.chs))
        return nn.Sequential(*blocks)

    def init_out_net(self, n_blocks):
        assert n_blocks >= 1
        chs = 4 * self.chs
        if self.cross:
            chs = 2 * self.chs
        blocks = []
        for _ in range(n_blocks - 1):
            blocks.append(self.block(chs))
        out_chs = 1
        if self.uncert:
            out_chs = 2
        elif self.discrete:
            out_chs = self.steps
        self.out_chs = out_chs
        blocks.append(self.block(chs, out_chs, False))
        return nn.Sequential(*blocks)

    def init_unet(self, depth=5):
        """feed_forward.py:189-204: the out-net as a U-Net (3x3 convs, BatchNorm, depth 5); head of 1 or 2 channels."""
        from .unet import UNet
        chs = (2 if self.cross else 4) * self.chs
        self.out_chs = 2 if self.uncert else 1
        return UNet(chs, self.out_chs, depth)

    # -- the engine, captured graphs and device tables are run-time caches: never copied or pickled with the module
    def __getstate__(self):
        state = dict(self.__dict__)
        state.update(_engine=None, _bins={}, _graphs={})
        return state

    # -- execution
    @property
    def engine(self):
        if self._engine is None:
            if self.ksize == 2 and not self.unet:
                self._engine = Engine(self)                  # published topology: tcgen05 implicit GEMM path
            else:
                from ..engine_generic import GenericEngine   # odd ksize / U-Net out-net: float32 layer kernels
                self._engine = GenericEngine(self)
        return self._engine

    def _bin_tables(self, device):
        key = str(device)
        if key not in self._bins:
            self._bins[key] = (ops.torch_bins(self.disp_min, self.disp_max, self.steps, device),
                               ops.numpy_bins(self.disp_min, self.disp_max, self.steps, device))
        return self._bins[key]

    def raw_forward(self, views, shift_disp=None):
        """(B, OC, H, W) network output without heads and without autograd (ESE inner loop)."""
        out, _ = self.engine.forward(views, self.training, save=False, shift_disp=shift_disp)
        return out

    # -- CUDA graphs for inference.  A full-light-field forward is ~45 launches of 20-130 us each: issued one by one
    # from Python the GPU waits for the host (measured: 3.5 ms per light field against 2.3 ms of kernel time), so the
    # eval-mode launch sequence is captured once per (input shapes, shifts, parameter version) and replayed.
    def _state_version(self):
        # version counters AND addresses: FusedAdam re-homes every p.data into its flat buffer without a version bump, and a
        # captured graph holds raw pointers (head weights, BatchNorm statistics)
        ts = list(self.parameters()) + list(self.buffers())
        return tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts) + (getattr(self, 'precision', 'fp16'),)

    def graphed_eval(self, views, shifts=(None,)):
        """Eval-mode raw network outputs for each entry of `shifts` (None = no Shift; a float = ESE member with the Shift
        fused into the packing), replayed from one CUDA graph.  Returns a list of (B, OC, H, W) tensors that stay valid
        until the next call with the same key."""
        assert not self.training
        key = (tuple(tuple(v.shape) for v in views), tuple(shifts), views[0].device.index)
        ver = self._state_version()
        hit = self._graphs.get(key)
        if hit is None or hit['ver'] != ver:
            static_in = [torch.empty_like(v) for v in views]
            for s, v in zip(static_in, views):
                s.copy_(v)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # eager warm-up: weight packs, BN folds, allocator pools
                self.engine.forward(static_in, False, save=False, shift_disp=shifts[0])
            torch.cuda.current_stream().wait_stream(side)
            from .. import _lib
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count
            with _lib.no_gc_during_capture(), torch.cuda.graph(graph):
                outs = [self.engine.forward(static_in, False, save=False, shift_disp=sd)[0] for sd in shifts]
            n_launches = _lib.launch_count - n0
            _lib.launch_count = n0                        # captured, not executed
            if len(self._graphs) >= 4:                    # bounded cache: drop the oldest capture (and its memory pool)
                self._graphs.pop(next(iter(self._graphs)))
            hit = {'ver': ver, 'graph': graph, 'in': static_in, 'outs': outs, 'launches': n_launches}
            self._graphs[key] = hit
        for s, v in zip(hit['in'], views):
            s.copy_(v)
        hit['graph'].replay()
        from .. import _lib
        _lib.launch_count += hit['launches']
        return hit['outs']

    def _can_graph(self, views):
        from .. import _lib
        return (self.use_cuda_graph and not self.training and _lib._profile is None and
                not torch.cuda.is_current_stream_capturing() and
                all(v.is_cuda and v.dtype == torch.float32 and v.is_contiguous() for v in views))

    def forward(self, h_views, v_views, i_views=None, d_views=None):
        """Same contract as feed_forward.py:206-305: stacks (b, n, 3, h, w) -> dict mean/logvar/scores/one_hot/posterior."""
        views = [h_views, v_views] if self.cross else [h_views, v_views, i_views, d_views]
        if not self.cross and (i_views is None or d_views is None):
            raise TypeError('the 4-stream model needs i_views and d_views (feed_forward.py:230-231)')
        params = [p for _, p in self.named_parameters()]
        # grad mode is off inside Function.forward, so decide here whether the backward tape is needed
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if not save and self._can_graph(views):
            output = self.graphed_eval(views)[0].clone()
        else:
            output = _NetFunction.apply(self.engine, self.training, save, len(views), *views, *params)
        mean = output[:, 0]
        eager = {'mean': mean, 'logvar': None, 'scores': None, 'one_hot': None, 'posterior': None}
        lazy = {}
        bins_t, bins_n = self._bin_tables(output.device)
        if self.discrete:
            # feed_forward.py:276-290.  mean/logvar/one_hot/posterior carry no gradient here; the reference only
            # trains DPP through `scores` (loss.py:145-160).
            scores = output
            det = scores.detach()

            def head():
                one_hot, post, m, lv = ops.dpp_head(det, bins_t, bins_n)
                return {'one_hot': one_hot, 'posterior': post, 'mean': m, 'logvar': lv}
            eager = {'scores': scores}
            lazy = {'mean': head, 'logvar': head, 'one_hot': head, 'posterior': head}
        if self.uncert:
            logvar = output[:, 1]                                        # feed_forward.py:293
            eager['logvar'] = logvar

            def post():
                return {'posterior': ops.upr_posterior(mean.detach(), logvar.detach(), bins_n)}
            lazy = {'posterior': post}
        out = LazyOutputs(eager, lazy)
        return out
