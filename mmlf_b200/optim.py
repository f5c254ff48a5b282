"""Fused Adam over one flat parameter buffer: one kernel launch per step instead of torch's foreach Adam over 82
tensors.  Same update rule and the same ``state_dict`` format as ``torch.optim.Adam`` (which is what the reference
constructs at /root/reference/mmlf/train/cli.py:113-118), so optimizer states in ``checkpoint.pt`` interchange."""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._flat = None

    def _flatten(self):
        """Re-home all parameters (and their Adam moments) as views of contiguous flat buffers."""
        ps = [p for g in self.param_groups for p in g['params']]
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            k = p.numel()
            flat_p[off:off + k].copy_(p.detach().reshape(-1))
            p.data = flat_p[off:off + k].view_as(p)
            st = self.state[p]
            if 'exp_avg' in st:                      # resumed from a checkpoint
                flat_m[off:off + k].copy_(st['exp_avg'].reshape(-1))
                flat_v[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
            else:
                st['step'] = torch.tensor(0.0)
            st['exp_avg'] = flat_m[off:off + k].view_as(p)
            st['exp_avg_sq'] = flat_v[off:off + k].view_as(p)
            if p.grad is not None:
                flat_g[off:off + k].copy_(p.grad.reshape(-1))
            p.grad = flat_g[off:off + k].view_as(p)
            off += k
        self._flat = (ps, flat_p, flat_g, flat_m, flat_v)

    def broadcast_state_(self, src=0):
        """Every rank takes rank `src`'s flat parameters and Adam moments (no-op single-process); see
        parallel.broadcast_module_."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if self._flat is None:
            self._flatten()
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        for t in (flat_p, flat_m, flat_v):
            dist.broadcast(t, src)
        steps = torch.tensor([float(self.state[ps[0]]['step'])], device=flat_p.device)
        dist.broadcast(steps, src)
        for p in ps:
            self.state[p]['step'] = torch.tensor(float(steps.item()))
        torch.autograd.graph.increment_version(ps)

    def _adopted(self):
        """The gradients as ONE flat buffer without copying, when every p.grad is a dense view, in parameter order, of
        one foreign allocation -- which is what Engine.backward hands to autograd after ``zero_grad()`` set the
        gradients to None (autograd adopts the views).  Returns that buffer as a 1-D tensor, or None."""
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        g0 = ps[0].grad
        if g0 is None or g0.data_ptr() == flat_g.data_ptr():
            return None
        base, off = g0.data_ptr(), 0
        for p in ps:
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * off:
                return None
            off += p.numel()
        st = g0.untyped_storage()
        if g0.storage_offset() * 4 + off * 4 > st.nbytes():
            return None
        return torch.empty(0, dtype=torch.float32, device=g0.device).set_(st, g0.storage_offset(), (off,), (1,))

    def _gather_grads(self):
        """-> the flat gradient buffer holding the current p.grad values (adopted in place, or copied where needed)."""
        if self._flat is None:
            self._flatten()
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        ad = self._adopted()
        if ad is not None:
            return ad
        off = 0
        for p in ps:                                  # gradients written elsewhere (autograd re-created .grad)
            k = p.numel()
            if p.grad is None:
                flat_g[off:off + k].zero_()
                p.grad = flat_g[off:off + k].view_as(p)
            elif p.grad.data_ptr() != flat_g.data_ptr() + off * 4:
                flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = flat_g[off:off + k].view_as(p)
            off += k
        return flat_g

    @property
    def flat_grad(self):
        """The flat gradient buffer (all-reduce this in data-parallel training)."""
        return self._gather_grads()

    @property
    def flat_buffers(self):
        """(params, grads, exp_avg, exp_avg_sq) flat float32 buffers owned by the optimizer, in parameter order."""
        if self._flat is None:
            self._flatten()
        return self._flat[1:]

    def zero_grad(self, set_to_none=True):
        """set_to_none=True (default, like torch): drop the gradients -- the next backward pass hands autograd views of
        the engine's flat gradient buffer, which `step` then uses in place (no per-parameter add / copy kernels).
        set_to_none=False: keep p.grad as views of the optimizer's own buffer, cleared with one memset."""
        if self._flat is None:
            self._flatten()
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        if set_to_none:
            for p in ps:
                p.grad = None
            return
        ops.zero_(flat_g)
        off = 0
        for p in ps:
            if p.grad is None or p.grad.data_ptr() != flat_g.data_ptr() + off * 4:
                p.grad = flat_g[off:off + p.numel()].view_as(p)
            off += p.numel()

    def host_step(self):
        """Current step count (host mirror of the optimizer state)."""
        if self._flat is None:
            self._flatten()
        st = self.state[self._flat[0][0]]['step']
        return int(st.item()) if isinstance(st, torch.Tensor) else int(st)

    def set_host_step(self, step):
        for p in self._flat[0]:
            self.state[p]['step'] = torch.tensor(float(step))

    @torch.no_grad()
    def step(self, closure=None):
        if self._flat is None:
            self._flatten()
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        grads = self._gather_grads()
        g0 = self.param_groups[0]
        step = self.host_step() + 1
        assert len(self.param_groups) == 1, 'FusedAdam keeps one flat buffer: use a single param group'
        ops.adam_step(flat_p, grads, flat_m, flat_v, float(g0['lr']), g0['betas'][0], g0['betas'][1], g0['eps'], step)
        self.set_host_step(step)
        torch.autograd.graph.increment_version(ps)    # the kernel wrote through raw pointers
        return None
