"""World-size-2 `gloo` tests (CPU) of the multi-rank host logic: batch sharding, the flat gradient bucket, the
loss-normaliser exchange and the ESE member sharding.

The product kernels need a B200, so the two ranks run the *host* code of `mmlf_b200` (parallel.py, model/loss.py,
model/ensamble.py) with the `ops.*` kernel wrappers replaced by stand-ins that follow the kernels' contract and are
built from the oracle.  What is checked is the exchange protocol: results on 2 ranks == results of one process on
the whole batch.
"""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, q):
    try:
        for p in (ROOT, os.path.join(ROOT, 'tests')):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.set_num_threads(1)
        from mmlf_b200 import parallel
        r, w, _ = parallel.init_from_env('gloo')
        assert (r, w) == (rank, world)
        out = globals()[fn_name](rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, 'ok', out))
    except Exception:  # pragma: no cover
        q.put((rank, 'error', traceback.format_exc()))


def _run(fn_name, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, status, out = q.get(timeout=180)
        assert status == 'ok', f'rank {rank}:\n{out}'
        res[rank] = out
    for p in procs:
        p.join(timeout=30)
    return res


# ------------------------------------------------------------------------------------------- stand-ins for ops.*
def _install_loss_standins():
    """CPU stand-ins with the contract of mmlf_loss_prepass / mmlf_loss_regression (include/mmlf_b200.h): the
    pre-pass returns LOCAL normalisers, the main pass takes the (all-reduced) GLOBAL ones and returns the local
    un-normalised loss sum plus gradients already divided by the global count."""
    from mmlf_b200 import ops

    def loss_prepass(mask, mask_padding=None, mpi=None):
        s = torch.zeros(8, dtype=torch.float64)
        s[0] = mask.sum()
        if mask_padding is not None:
            s[1] = mask_padding.sum()
        if mpi is not None:
            w = mpi[:, :, 3].double().sum(1)
            s[2] = w.sum()
            s[3] = (mpi[:, :, 3].sum(1) < 0.01).sum()
        s[4] = mask.numel()
        return s

    def loss_regression(kind, mean, logvar, target, mask, mask_padding, sums, param=0.0, want_grad=True):
        cnt, N = sums[0].item(), sums[4].item()
        m = mask.double()
        mean_d, d = mean.double(), mean.double() - target.double() if kind in (0, 2, 4, 5) else None
        if kind == 0:
            per, gm, gl = d.abs(), d.sign(), None
        elif kind == 2:
            lv = logvar.double()
            e = torch.exp(-lv)
            per, gm, gl = e * d.abs() + lv, e * d.sign(), -e * d.abs() + 1.0
            if mask_padding is not None:
                mp_ = mask_padding.double()
                k_in, k_oor = N / sums[1].item(), N / (N - sums[1].item())
                per = (per * mp_ * k_in - lv * (1 - mp_) * k_oor) / 2
                gm = gm * mp_ * k_in / 2
                gl = (gl * mp_ * k_in - (1 - mp_) * k_oor) / 2
        elif kind == 3:
            lv = logvar.double()
            e = torch.exp(-lv)
            w, t = target[:, :, 3].double(), target[:, :, 4].double()
            dd = mean_d[:, None] - t
            mw = sums[2].item() / N
            k = N / sums[3].item()
            oor = (target[:, :, 3].sum(1) < 0.01).double()
            per = (((e[:, None] * dd.abs() + lv[:, None]) * w).sum(1) / mw - lv * oor * k) / 2
            gm = (e[:, None] * dd.sign() * w).sum(1) / mw / 2
            gl = (((-e[:, None] * dd.abs() + 1.0) * w).sum(1) / mw - oor * k) / 2
        else:
            raise NotImplementedError(kind)
        scale = 1.0 if cnt == 0 else 1.0 / cnt
        loss_sum = (per * m).sum().reshape(1)
        g_mean = (gm * m * scale).float() if want_grad else None
        g_logvar = (gl * m * scale).float() if (want_grad and gl is not None) else None
        return loss_sum, g_mean, g_logvar

    def loss_finish(loss_sum, sums):
        cnt = sums[0]
        return (loss_sum[0] / torch.where(cnt == 0, torch.ones_like(cnt), cnt)).to(torch.float32)

    ops.loss_prepass = loss_prepass
    ops.loss_regression = loss_regression
    ops.loss_finish = loss_finish


def _loss_inputs():
    rng = np.random.RandomState(3)
    B, H, W, K = 6, 12, 10, 3
    mean = rng.normal(0, 1, (B, H, W)).astype(np.float32)
    logvar = rng.normal(0, 0.5, (B, H, W)).astype(np.float32)
    gt = rng.normal(0, 1, (B, H, W)).astype(np.float32)
    mask = (rng.uniform(size=(B, H, W)) > 0.3)
    mask[:2] &= rng.uniform(size=(2, H, W)) > 0.6          # uneven mask counts per rank
    pad = (rng.uniform(size=(B, H, W)) > 0.2)
    mpi = rng.uniform(0, 1, (B, K, 5, H, W)).astype(np.float32)
    mpi[:, :, 3] *= (rng.uniform(size=(B, 1, H, W)) > 0.1)   # some pixels without any plane (OOR)
    mpi[:, :, 4] = rng.normal(0, 1, (B, K, H, W))
    return mean, logvar, gt, mask, pad, mpi


# ------------------------------------------------------------------------------------------- rank bodies
def _body_losses(rank, world):
    import oracle
    from mmlf_b200 import parallel
    from mmlf_b200.model import loss as L
    _install_loss_standins()
    mean, logvar, gt, mask, pad, mpi = _loss_inputs()
    T = torch.from_numpy
    full = [T(mean), T(logvar), T(gt), T(mask), T(pad), T(mpi)]
    smean, slogvar, sgt, smask, spad, smpi = parallel.shard_batch(full, rank, world)
    lo = rank * ((mean.shape[0] + world - 1) // world)
    hi = lo + smean.shape[0]
    report = {}
    cases = {
        'l1': (L.MaskedL1Loss(), lambda o: (o, sgt, smask), lambda: oracle.masked_l1({'mean': mean}, gt, mask)),
        'upr': (L.ImprovedUncertaintyL1Loss(), lambda o: (o, sgt, smask),
                lambda: oracle.improved_uncertainty_l1({'mean': mean, 'logvar': logvar}, gt, mask)),
        'upr_pad': (L.ImprovedUncertaintyL1Loss(), lambda o: (o, sgt, smask, spad),
                    lambda: oracle.improved_uncertainty_l1({'mean': mean, 'logvar': logvar}, gt, mask, pad)),
        'multi_upr': (L.ImprovedMultiUncertaintyL1Loss(), lambda o: (o, smpi, smask),
                      lambda: oracle.improved_multi_uncertainty_l1({'mean': mean, 'logvar': logvar}, mpi, mask)),
    }
    for name, (mod, args, ref) in cases.items():
        m_ = smean.clone().requires_grad_(True)
        l_ = slogvar.clone().requires_grad_(True)
        val = mod(*args({'mean': m_, 'logvar': l_}))
        val.backward()
        rv, rg = ref()
        assert abs(val.item() - rv) < 1e-5 * max(1.0, abs(rv)), (name, val.item(), rv)
        assert np.allclose(m_.grad.numpy(), rg['mean'][lo:hi], rtol=1e-5, atol=1e-7), name
        if 'logvar' in rg:
            assert np.allclose(l_.grad.numpy(), rg['logvar'][lo:hi], rtol=1e-5, atol=1e-7), name
        report[name] = val.item()
    return report


def _body_bucket(rank, world):
    """Sharded batch + flat bucket all-reduce == full-batch gradient (sum-reduced loss, like the kernels' gradients,
    which already carry the 1/global-count factor)."""
    from mmlf_b200 import parallel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 2, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(5, 1, 2))
    x = torch.randn(8, 3, 9, 9)
    y = torch.randn(8, 1, 9, 9)
    ref = [g.clone() for g in torch.autograd.grad(((net(x) - y) ** 2).sum() / x.shape[0], list(net.parameters()))]
    bucket = parallel.GradBucket(net.parameters())
    bucket.zero_()
    xs, ys = parallel.shard_batch([x, y], rank, world)
    (((net(xs) - ys) ** 2).sum() / x.shape[0]).backward()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    bucket.all_reduce()
    for p, r in zip(net.parameters(), ref):
        assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-6)
    # uneven shard: 7 samples over 2 ranks -> 4 + 3
    a, = parallel.shard_batch([torch.arange(7)], rank, world)
    return a.tolist()


class _FakeNet(torch.nn.Module):
    """Stand-in for FeedForward.raw_forward in eval mode: a deterministic function of the views and the shift."""
    cross = False

    def raw_forward(self, feed, shift_disp=0.0):
        self.calls.append(shift_disp)
        base = feed[0][:, 4, 0] + 0.5 * feed[3][:, 2, 1]
        mean = base * (1.0 + 0.01 * shift_disp) - shift_disp
        logvar = torch.sin(base * 7.0 + shift_disp * 3.0)
        return torch.stack([mean, logvar], 1)


def _body_ese(rank, world):
    import oracle
    from mmlf_b200 import ops
    from mmlf_b200.model.ensamble import Ensamble

    def ese_reduce(means, logvars, disp):
        m, lv, post = oracle.ensemble_reduce(means.numpy(), logvars.numpy(), float(disp[0]), float(disp[-1]))
        return [torch.from_numpy(m), torch.from_numpy(lv), torch.from_numpy(post)]
    ops.ese_reduce = ese_reduce
    rng = np.random.RandomState(1)
    views = [torch.from_numpy(rng.uniform(0, 1, (1, 9, 3, 5, 7)).astype(np.float32)) for _ in range(4)]   # 35 pixels: uneven pixel shards (18 + 17)
    net = _FakeNet()
    net.calls = []
    ens = Ensamble(net, -1.0, 1.0, 0.3)                     # 7 members: ranks get 4 + 3
    out = ens(*views)
    shifts = [float(s) for s in np.arange(-1.0, 1.0, 0.3)]
    assert net.calls == shifts[rank::world], (net.calls, shifts)
    # single-process result
    ref_net = _FakeNet()
    ref_net.calls = []
    means = torch.stack([ref_net.raw_forward(views, s)[:, 0] + s for s in shifts])
    logvars = torch.stack([ref_net.raw_forward(views, s)[:, 1] for s in shifts])
    assert torch.equal(out['means'], means) and torch.equal(out['logvars'], logvars)
    rm, rl, rp = oracle.ensemble_reduce(means.numpy(), logvars.numpy(), -1.0, 1.0)
    assert np.array_equal(out['mean'].numpy(), rm) and np.array_equal(out['posterior'].numpy(), rp)
    return len(net.calls)


def _body_bands(rank, world):
    """Row-band sharding of full-image inference: band + halo rows in, own rows out; stitched == whole image."""
    from mmlf_b200 import parallel
    H, W, radius = 37, 8, 3
    lo, hi, a, b = parallel.band_rows(H, rank, world, radius)
    img = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W)
    # receptive field of radius 3: a (2r+1)-row box filter with zero padding
    k = torch.ones(1, 1, 2 * radius + 1, 1)
    whole = torch.nn.functional.conv2d(img, k, padding=(radius, 0))
    band = torch.nn.functional.conv2d(img[:, :, a:b], k, padding=(radius, 0))[:, :, lo - a:hi - a]
    out = parallel.gather_bands(band.contiguous(), H, rank, world, radius, dim=2)
    assert torch.equal(out, whole)
    return (lo, hi, a, b)


# ------------------------------------------------------------------------------------------- tests
def test_loss_normalisers_are_global_across_ranks():
    res = _run('_body_losses')
    assert res[0] == res[1]                                  # every rank reports the global loss value


def test_grad_bucket_allreduce_matches_full_batch():
    res = _run('_body_bucket')
    assert res[0] == [0, 1, 2, 3] and res[1] == [4, 5, 6]


def test_ese_members_round_robin_and_gather():
    res = _run('_body_ese')
    assert res == {0: 4, 1: 3}


def test_inference_row_bands_with_halo():
    res = _run('_body_bands')
    assert res[0][0] == 0 and res[0][1] == res[1][0] and res[1][1] == 37


@pytest.mark.parametrize('H,world,radius', [(512, 8, 11), (512, 3, 11), (40, 4, 11), (5, 8, 2)])
def test_band_rows_cover_image(H, world, radius):
    from mmlf_b200 import parallel
    rows = []
    for r in range(world):
        lo, hi, a, b = parallel.band_rows(H, r, world, radius)
        assert 0 <= a <= lo <= hi <= b <= H
        assert a == max(0, lo - radius) and b == min(H, hi + radius)
        rows += list(range(lo, hi))
    assert rows == list(range(H))


# ------------------------------------------------------------------------------------------- replica consistency
def _case_broadcast(rank, world):
    """every rank builds its model from its own RNG; parallel.broadcast_module_ and
    FusedAdam.broadcast_state_ make them one replica (parameters, buffers, Adam moments, step count)."""
    import _fixtures as fx
    from mmlf_b200 import parallel
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.optim import FusedAdam
    torch.manual_seed(100 + rank)                                  # different weights per rank
    m = FeedForward(**fx.model_kwargs('upr', False, chs=8))
    with torch.no_grad():
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.add_(float(rank))
    opt = FusedAdam(m.parameters(), lr=1e-3)
    opt._flatten()
    opt._flat[3].fill_(float(rank + 1))                            # exp_avg
    opt.set_host_step(7 * (rank + 1))
    before = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
    parallel.broadcast_module_(m)
    opt.broadcast_state_()
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()] +
                     [b.detach().reshape(-1).double().float() for b in m.buffers()] + [opt._flat[3], opt._flat[4]])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    same = all(torch.equal(parts[0], p) for p in parts)
    changed = bool((before != torch.cat([p.detach().reshape(-1) for p in m.parameters()])).any())
    return dict(same=same, changed=changed, step=opt.host_step(), views_ok=all(
        p.data_ptr() == opt._flat[1].data_ptr() + 4 * off for p, off in zip(
            m.parameters(), np.cumsum([0] + [q.numel() for q in m.parameters()][:-1]).tolist())))


def test_ranks_start_from_one_replica():
    res = _run('_case_broadcast')
    assert res[0]['same'] and res[1]['same']
    assert not res[0]['changed'] and res[1]['changed']              # rank 1 took rank 0's weights
    assert res[0]['step'] == res[1]['step'] == 7
    assert res[0]['views_ok'] and res[1]['views_ok']                # parameters are still views of the flat buffer
