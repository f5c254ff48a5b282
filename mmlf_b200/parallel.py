"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the only real exchange steps of the
path -- the gradient all-reduce, the loss normalisers and the ESE member gather.  Replaces the
single-process ``torch.nn.DataParallel`` of /root/reference/mmlf/train/cli.py:159."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend=backend)
    return rank, world, local


def shard_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_reduce_sum_(t):
    """In-place SUM all-reduce when running multi-rank; no-op otherwise."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_module_(module, src=0):
    """Every rank takes rank `src`'s parameters and buffers (no-op single-process).  The data-parallel training loop
    only exchanges gradients, so the replicas must START identical -- torch.nn.DataParallel gets this by re-replicating
    from device 0 every step (train/cli.py:159)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return module
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
    torch.autograd.graph.increment_version([t for t in list(module.parameters()) + list(module.buffers())])
    return module


def shard_batch(tensors, rank, world):
    """Contiguous dim-0 shard of each tensor (DataParallel scatter, train/cli.py:245)."""
    out = []
    for t in tensors:
        n = t.shape[0]
        per = (n + world - 1) // world
        out.append(t[rank * per:min(n, (rank + 1) * per)])
    return out


def gather_members(means, logvars, K, rank, world):
    """ESE: every rank computed members rank, rank + world, ...; exchange them so each rank holds all K.  One all-gather
    of the ranks' (padded) member stacks instead of 2 K broadcasts of one plane each (140 launches of ~1 MB for the
    70-member ensemble, latency bound)."""
    n_max = (K + world - 1) // world
    n_mine = len(range(rank, K, world))
    buf = means.new_zeros((2, n_max) + tuple(means.shape[1:]))
    buf[0, :n_mine] = means[rank::world]
    buf[1, :n_mine] = logvars[rank::world]
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    for r in range(world):
        if r == rank:
            continue
        n_r = len(range(r, K, world))
        means[r::world] = parts[r][0, :n_r]
        logvars[r::world] = parts[r][1, :n_r]


def sharded_ese_reduce(reduce_fn, means, logvars, disp, rank, world):
    """The min-logvar pick and the K x K Laplace mixture (ensamble.py:78-101) are independent per pixel: each rank
    reduces its 1 / world share of the pixels and the shares are all-gathered (the 70-member reduce is 2.5 ms of SFU
    work per 512 x 512 light field -- replicated on every rank it was 8 % of the 8-GPU ESE step).  ``reduce_fn`` is
    ``ops.ese_reduce``; batches (B > 1) and single-process runs take the plain call."""
    K, B, H, W = means.shape
    HW = H * W
    if world == 1 or B != 1:
        return reduce_fn(means, logvars, disp)
    chunk = (HW + world - 1) // world
    p0 = min(rank * chunk, HW)
    p1 = min(p0 + chunk, HW)
    m_s = means.new_zeros((K, 1, 1, chunk))
    l_s = means.new_zeros((K, 1, 1, chunk))
    m_s[:, 0, 0, :p1 - p0] = means.reshape(K, HW)[:, p0:p1]
    l_s[:, 0, 0, :p1 - p0] = logvars.reshape(K, HW)[:, p0:p1]
    mean_s, logvar_s, post_s = reduce_fn(m_s, l_s, disp)
    pack = torch.cat([mean_s.reshape(1, chunk), logvar_s.reshape(1, chunk), post_s.reshape(K, chunk)], 0).contiguous()
    parts = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(parts, pack)
    full = torch.cat(parts, 1)[:, :HW]                                 # (K + 2, HW)
    return [full[0].reshape(B, H, W).contiguous(), full[1].reshape(B, H, W).contiguous(),
            full[2:].reshape(B, K, H, W).contiguous()]


class GradBucket:
    """All parameter gradients of a module in one flat fp32 buffer -> a single all-reduce per step
    (18.4 MB for the 4-stream model) instead of DataParallel's replicate/scatter/gather/reduce."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()
        off = 0
        for p in self.params:          # re-attach in case an optimizer set .grad to None
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * 4:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce(self):
        all_reduce_sum_(self.flat)


# ------------------------------------------------------------------------------------------------- full-image bands
def band_rows(H, rank, world, radius):
    """Row band of rank `rank` when an H-row image is split over `world` ranks: returns
    (lo, hi, a, b) -- the rank owns output rows [lo, hi) and reads input rows [a, b) = the band plus a `radius`-row
    halo clipped to the image (at the image border the network's own zero padding applies, which is exact)."""
    per = (H + world - 1) // world
    lo = min(H, rank * per)
    hi = min(H, lo + per)
    return lo, hi, max(0, lo - radius), min(H, hi + radius)


def gather_bands(band, H, rank, world, radius, dim):
    """All-gather the owned rows of every rank along `dim` -> the whole image on every rank."""
    if world == 1:
        return band
    per = (H + world - 1) // world
    shape = list(band.shape)
    shape[dim] = per
    padded = band.new_zeros(shape)
    padded.narrow(dim, 0, band.shape[dim]).copy_(band)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    rows = [p.narrow(dim, 0, band_rows(H, r, world, radius)[1] - band_rows(H, r, world, radius)[0])
            for r, p in enumerate(parts)]
    return torch.cat(rows, dim)


def banded_forward(model, views, radius=11, full_height=None):
    """Full-image inference of one light field sharded by rows over the ranks: the FCN's receptive field is
    `radius` = 11 px for the published topology (3 + 8 blocks of 2x2 conv pairs; `model_radius` in train/cli.py:95),
    so each rank runs the network on its band + halo and keeps its own rows.  Any circular Shift must already have
    been applied to the whole image.  With `full_height` the caller passes only its band's rows.  Returns the gathered dict {'mean', 'logvar'} (B, H, W) / {'scores'} (B, S, H, W)."""
    rank, world = shard_info()
    if full_height is None:
        H = views[0].shape[-2]
        lo, hi, a, b = band_rows(H, rank, world, radius)
        out = model(*[None if v is None else v[..., a:b, :].contiguous() for v in views])
    else:
        # `views` already hold only this rank's rows [a, b) of a light field of `full_height` rows (a loader that reads or
        # uploads just its band: band_rows(full_height, rank, world, radius) says which)
        H = full_height
        lo, hi, a, b = band_rows(H, rank, world, radius)
        assert views[0].shape[-2] == b - a, (views[0].shape, a, b)
        out = model(*views)
    res = {}
    for k in ('mean', 'logvar', 'scores'):
        t = out[k]
        if t is not None:
            res[k] = gather_bands(t[..., lo - a:hi - a, :].contiguous(), H, rank, world, radius, dim=t.dim() - 2)
    return res
