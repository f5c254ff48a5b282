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

    @property
    def flat_grad(self):
        """The flat gradient buffer (all-reduce this in data-parallel training)."""
        if self._flat is None:
            self._flatten()
        return self._flat[2]

    def zero_grad(self, set_to_none=False):
        if self._flat is None:
            self._flatten()
        self._flat[2].zero_()

    @torch.no_grad()
    def step(self, closure=None):
        if self._flat is None:
            self._flatten()
        ps, flat_p, flat_g, flat_m, flat_v = self._flat
        off = 0
        for p in ps:                                  # gradients written elsewhere (autograd re-created .grad)
            k = p.numel()
            if p.grad is not None and p.grad.data_ptr() != flat_g.data_ptr() + off * 4:
                flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = flat_g[off:off + k].view_as(p)
            off += k
        g0 = self.param_groups[0]
        step = int(self.state[ps[0]]['step'].item()) + 1 if isinstance(self.state[ps[0]]['step'], torch.Tensor) \
            else int(self.state[ps[0]]['step']) + 1
        assert len(self.param_groups) == 1, 'FusedAdam keeps one flat buffer: use a single param group'
        ops.adam_step(flat_p, flat_g, flat_m, flat_v, float(g0['lr']), g0['betas'][0], g0['betas'][1], g0['eps'], step)
        for p in ps:
            self.state[p]['step'] = torch.tensor(float(step))
        torch.autograd.graph.increment_version(ps)    # the kernel wrote through raw pointers
        return None
