"""One training step of the reference loop (/root/reference/mmlf/train/cli.py:243-258: forward, loss, backward,
``optimizer.step()``) as ONE replayable CUDA graph of hand-written kernels:

    pack views -> conv / BatchNorm chain -> head -> loss pre-pass [-> all-reduce of the normalisers] -> loss value +
    gradient -> hand-derived backward chain (gradients land straight in the optimizer's flat buffer) [-> all-reduce of
    the flat gradient bucket over NCCL] -> Adam

No autograd, no framework compute kernel and -- once captured -- no per-launch host work: at 64 patches per GPU the
host needed ~20 ms to enqueue a 27 ms step, which capped the 8-GPU scaling.  Everything that changes
between replays lives in device memory: the inputs (static buffers the loader writes into), the learning rate and the
Adam step count (``mmlf_adam_step_dev``), BatchNorm's ``num_batches_tracked``.

The eager autograd path (``loss_fn(model(...)).backward(); optimizer.step()``) stays available and is what the parity
tests compare this against; both run the same kernels through ``Engine.forward`` / ``Engine.backward``.
"""
import ctypes as C

import torch

from .. import _lib, ops, parallel
from .._lib import call

# loss name -> (kind of mmlf_loss_regression or 'ce', needs logvar, multi-plane target)
LOSSES = {
    'l1': (0, False, False),            # MaskedL1Loss                      (loss.py:46-77)
    'multi_l1': (1, False, True),       # MultiMaskedL1Loss                 (loss.py:88-103)
    'upr': (2, True, False),            # ImprovedUncertaintyL1Loss         (loss.py:262-294)
    'multi_upr': (3, True, True),       # ImprovedMultiUncertaintyL1Loss    (loss.py:344-372)
    'ce': ('ce', False, False),         # MaskedCrossEntropy                (loss.py:145-160)
}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class TrainStep:
    """``step = TrainStep(model, optimizer, 'upr'); loss = step(h, v, i, d, target, mask[, mask_padding])``.

    model: FeedForward (``model.training`` decides train- or eval-mode BatchNorm, as train/cli.py:227-230);
    optimizer: mmlf_b200.optim.FusedAdam over ``model.parameters()``; loss: a key of LOSSES.  For 'ce' the target is
    either the dense (B, S, H, W) class target or, with ``ce_from_gt=True``, the (B, H, W) disparity from which the
    kernel builds the one-hot target on the fly (utils/dl.py:109-131).
    ``lr`` is read from ``optimizer.param_groups[0]['lr']`` at every call, so warm start / cooling work unchanged.
    Returns the loss as a 0-d float32 CUDA tensor that is overwritten by the next call (read it with ``.item()`` only
    when you need it: that is the step's one host sync)."""

    def __init__(self, model, optimizer, loss, ce_from_gt=False, use_graph=True):
        if loss not in LOSSES:
            raise ValueError(f'unknown loss {loss!r}; one of {sorted(LOSSES)}')
        self.model, self.opt, self.loss = model, optimizer, loss
        self.kind, self.needs_logvar, self.multi = LOSSES[loss]
        if self.needs_logvar and not model.uncert:
            raise ValueError(f"loss {loss!r} needs a --model_uncert network")
        if self.kind == 'ce' and not model.discrete:
            raise ValueError("loss 'ce' needs a --model_discrete network")
        self.ce_from_gt = ce_from_gt
        self.use_graph = use_graph
        self._graphs = {}
        self._hyper_host = None
        self.launches_per_step = 0
        self.replays = 0

    def close(self):
        """Drop the captured graphs (and the activation memory they own).  Call it before tearing down a multi-rank
        process group: a live graph holds NCCL kernels of that communicator, and destroying the group under it can hang."""
        import gc
        torch.cuda.synchronize()
        self._graphs.clear()
        gc.collect()
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the work of one step, on static buffers
    def _body(self, S):
        m, eng = self.model, self.model.engine
        views = S['views']
        flat_p, flat_g, flat_m, flat_v = self.opt.flat_buffers
        # the packed tensor-core operands are re-derived from the fp32 parameters in EVERY step (one launch): the step
        # itself changes them, and a captured step must contain that launch whatever the version counters say
        eng._pack_version = None
        out, tape = eng.forward(views, m.training, save=True)
        B, OC, H, W = out.shape
        HW = H * W
        mask, mp, tgt = S['mask'], S.get('mask_padding'), S['target']
        sums = S['sums']
        ops.zero_(S['acc'])                                   # sums (8 doubles) + loss_sum (1): one memset
        loss_sum = S['loss_sum']
        call('mmlf_loss_prepass', _p(mask), _p(mp), _p(tgt if self.multi else None), tgt.shape[1] if self.multi else 0,
             B, HW, _p(sums), _st())
        parallel.all_reduce_sum_(sums)                        # global normalisers BEFORE the gradient scale is fixed
        g_out = S['g_out']
        if self.kind == 'ce':
            dense = None if self.ce_from_gt else tgt
            gt = tgt if self.ce_from_gt else None
            bins = m._bin_tables(out.device)[0] if self.ce_from_gt else None
            half = (m.disp_max - m.disp_min) / m.steps / 2.0
            call('mmlf_loss_cross_entropy', _p(out), _p(dense), _p(gt), _p(bins), float(half), OC, _p(mask), _p(sums), B,
                 HW, _p(loss_sum), _p(g_out), _st())
        else:
            # mean / logvar are planes 0 / 1 of the network output and their gradients planes of g_out: in place
            lv = C.c_void_p(out.data_ptr() + 4 * HW) if self.needs_logvar else C.c_void_p(0)
            glv = C.c_void_p(g_out.data_ptr() + 4 * HW) if self.needs_logvar else C.c_void_p(0)
            call('mmlf_loss_regression', self.kind, _p(out), lv, _p(tgt), tgt.shape[1] if self.multi else 0, _p(mask),
                 _p(mp), _p(sums), 0.0, B, HW, _p(loss_sum), _p(g_out), glv, OC * HW, _st())
        parallel.all_reduce_sum_(loss_sum)
        call('mmlf_loss_finish', _p(loss_sum), _p(sums), _p(S['loss']), _st())
        ops.zero_(flat_g)
        eng.backward(tape, g_out, flat=flat_g)
        parallel.all_reduce_sum_(flat_g)                      # replaces DataParallel's reduce-add (train/cli.py:159)
        g0 = self.opt.param_groups[0]
        S['hyper'][:1].copy_(self._hyper_host[:1], non_blocking=True)     # lr: pinned host -> device (a memcpy node)
        call('mmlf_adam_step_dev', _p(flat_p), _p(flat_g), _p(flat_m), _p(flat_v), flat_p.numel(), _p(S['hyper']),
             g0['betas'][0], g0['betas'][1], g0['eps'], _st())

    def _static(self, views, target, mask, mask_padding):
        dev = views[0].device
        B, n, c3, H, W = views[0].shape
        acc = torch.empty(9, dtype=torch.float64, device=dev)
        S = dict(views=[torch.empty_like(v) for v in views], target=torch.empty_like(target),
                 mask=torch.empty_like(mask), acc=acc, sums=acc[:8], loss_sum=acc[8:9],
                 loss=torch.empty(1, dtype=torch.float32, device=dev),
                 g_out=torch.empty((B, self.model.out_chs, H, W), dtype=torch.float32, device=dev),
                 hyper=ops.zero_(torch.empty(4, dtype=torch.float64, device=dev)))
        if mask_padding is not None:
            S['mask_padding'] = torch.empty_like(mask_padding)
        if self.kind != 'ce' and self.model.out_chs > 2:
            raise ValueError('regression losses need a BASE / UPR head')
        if not self.needs_logvar and self.kind != 'ce' and self.model.out_chs == 2:
            ops.zero_(S['g_out'])                             # UPR network trained with a plain L1 loss: d/dlogvar = 0
        return S

    @staticmethod
    def _norm(views, target, mask, mask_padding):
        views = [v for v in views if v is not None]
        for v in views:
            if not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
                raise RuntimeError('TrainStep: view stacks must be contiguous float32 CUDA tensors')
        target = target if target.dtype == torch.float32 else target.float()
        mask = mask if mask.dtype == torch.int32 else mask.to(torch.int32)
        if mask_padding is not None and mask_padding.dtype != torch.int32:
            mask_padding = mask_padding.to(torch.int32)
        return views, target.contiguous(), mask.contiguous(), None if mask_padding is None else mask_padding.contiguous()

    # ------------------------------------------------------------------ call
    def buffers(self, views, target, mask, mask_padding=None):
        """The static input buffers for inputs of these shapes (views list, target, mask, mask_padding): a loader may
        write the next batch straight into them and then call ``step()`` without arguments' copies."""
        views, target, mask, mask_padding = self._norm(views, target, mask, mask_padding)
        S = self._get(views, target, mask, mask_padding)
        return S['views'], S['target'], S['mask'], S.get('mask_padding')

    def _key(self, views, target, mask, mask_padding):
        return (tuple(views[0].shape), len(views), tuple(target.shape), mask_padding is not None, self.model.training,
                views[0].device.index, getattr(self.model, 'precision', 'fp16'))

    def _get(self, views, target, mask, mask_padding):
        key = self._key(views, target, mask, mask_padding)
        S = self._graphs.get(key)
        if S is None:
            if self._hyper_host is None:
                self._hyper_host = torch.zeros(2, dtype=torch.float64).pin_memory()      # [lr, step count]
            self.opt.flat_buffers                              # flatten parameters before anything captures pointers
            S = self._static(views, target, mask, mask_padding)
            S['graph'] = None
            if len(self._graphs) >= 2:                        # each capture owns the activation memory of a whole step
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = S
        return S

    def __call__(self, h_views, v_views, i_views, d_views, target, mask, mask_padding=None):
        _lib.require_device()
        stacks = [h_views, v_views] if self.model.cross else [h_views, v_views, i_views, d_views]
        views, target, mask, mask_padding = self._norm(stacks, target, mask, mask_padding)
        S = self._get(views, target, mask, mask_padding)
        for dst, src in zip(S['views'] + [S['target'], S['mask']], views + [target, mask]):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if mask_padding is not None and S['mask_padding'].data_ptr() != mask_padding.data_ptr():
            S['mask_padding'].copy_(mask_padding, non_blocking=True)
        return self.step(S)

    def step(self, S=None):
        """Run one step on the static buffers (the last shapes used when ``S`` is None)."""
        if S is None:
            S = next(reversed(self._graphs.values()))
        opt = self.opt
        self._hyper_host[0] = float(opt.param_groups[0]['lr'])
        step0 = opt.host_step()
        params = list(self.model.parameters())
        buffers = list(self.model.buffers())
        capturing = torch.cuda.is_current_stream_capturing()
        if S.get('dev_step') != step0 and not capturing:
            # the device-side step counter advances by itself; re-seed it only when the host count moved under it
            # (first use, optimizer.load_state_dict, eager optimizer.step() calls in between)
            self._hyper_host[1] = float(step0)
            S['hyper'][1:2].copy_(self._hyper_host[1:2], non_blocking=True)
            torch.cuda.current_stream().synchronize()         # the pinned value may be rewritten by the next call
        S['dev_step'] = step0 + 1
        if not self.use_graph or capturing or _lib._profile is not None:
            n0 = _lib.launch_count
            self._body(S)
            self.launches_per_step = _lib.launch_count - n0
        else:
            if S['graph'] is None:
                # eager warm-up (NCCL communicators, weight-pack job tables, vec-job table, allocator), then capture.
                # The warm-up IS a training step; the capture itself executes nothing.
                n0 = _lib.launch_count
                self._body(S)
                self.launches_per_step = _lib.launch_count - n0
                torch.cuda.synchronize()
                torch.cuda.empty_cache()                      # the warm-up's activations must not double the footprint
                g = torch.cuda.CUDAGraph()
                n1 = _lib.launch_count
                with _lib.no_gc_during_capture(), torch.cuda.graph(g, capture_error_mode='thread_local'):
                    self._body(S)
                _lib.launch_count = n1
                S['graph'] = g
            else:
                S['graph'].replay()
                _lib.launch_count += self.launches_per_step
                self.replays += 1
        opt.set_host_step(step0 + 1)
        # the kernels wrote through raw pointers: version counters key the weight-pack / BN-fold / eval-graph caches
        torch.autograd.graph.increment_version(params + buffers)
        return S['loss'][0]
