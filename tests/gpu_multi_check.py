"""N-GPU correctness check (torchrun, one rank per GPU): the member-sharded / pixel-sharded ESE and the row-band sharded
inference give the same result as the single-process path on the same light field.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_multi_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmlf_b200 import parallel  # noqa: E402
from mmlf_b200.model.ensamble import Ensamble  # noqa: E402
from mmlf_b200.model.feed_forward import FeedForward  # noqa: E402

KW = dict(model_ksize=2, model_in_blocks=3, model_out_blocks=8, model_chs=70, model_views=9, model_cross=False,
          model_uncert=True, model_unet=False, model_discrete=False, model_no_batchnorm=False,
          model_batchnorm_momentum=0.1, val_disp_min=-3.5, val_disp_max=3.5)


def main():
    parallel.init_from_env()
    rank, world = parallel.shard_info()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = FeedForward(**KW).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(1234)
    H = W = 200                                                    # 40000 pixels: not a multiple of 64 * world
    views = [torch.rand((1, 9, 3, H, W), device=dev, generator=gen) for _ in range(4)]
    ens = Ensamble(model, -3.5, 3.5, 0.5)                          # 14 members
    with torch.no_grad():
        out = ens(*views)
        band = parallel.banded_forward(model, views)
        # single-process reference on this rank
        saved = parallel.shard_info
        parallel.shard_info = lambda: (0, 1)
        try:
            ref = ens(*views)
            whole = model(*views)
        finally:
            parallel.shard_info = saved
    ok = True
    for k in ('mean', 'logvar', 'means', 'logvars', 'posterior'):
        same = torch.equal(out[k], ref[k])
        err = (out[k] - ref[k]).abs().max().item()
        print(f'rank {rank} ese {k}: equal={same} max_abs_diff={err:.3e}')
        ok &= err <= 1e-6
    for k in ('mean', 'logvar'):
        err = (band[k] - whole[k]).abs().max().item()
        print(f'rank {rank} bands {k}: max_abs_diff={err:.3e}')
        ok &= err <= 1e-6
    # ---- data-parallel training: ranks built from DIFFERENT seeds become one replica after the rank-0 broadcast and stay
    # bit-identical over captured TrainStep steps on different data (NCCL all-reduce of normalisers + gradients in-graph)
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep
    torch.manual_seed(100 + rank)
    tm = FeedForward(**dict(KW, model_chs=16)).to(dev).train()
    opt = FusedAdam(tm.parameters(), lr=1e-3)
    parallel.broadcast_module_(tm)
    opt.broadcast_state_()
    step = TrainStep(tm, opt, 'upr')
    g2 = torch.Generator(device=dev).manual_seed(77 + rank)        # every rank trains on its own shard
    losses = []
    for it in range(4):
        tv = [torch.rand((4, 9, 3, 32, 32), device=dev, generator=g2) for _ in range(4)]
        tgt = torch.rand((4, 32, 32), device=dev, generator=g2) * 2 - 1
        tmask = (torch.rand((4, 32, 32), device=dev, generator=g2) > 0.2).to(torch.int32)
        losses.append(step(*tv, tgt, tmask).item())
    flat = opt.flat_buffers[0]
    parts = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(parts, flat)
    same = all(torch.equal(parts[0], p) for p in parts)
    lt = torch.tensor(losses, device=dev, dtype=torch.float64)
    lparts = [torch.empty_like(lt) for _ in range(world)]
    torch.distributed.all_gather(lparts, lt)
    lsame = all(torch.equal(lparts[0], p) for p in lparts)          # the loss is the GLOBAL masked mean on every rank
    print(f'rank {rank} train replicas: params identical={same} losses identical={lsame} replays={step.replays} '
          f'loss {losses[0]:.5f} -> {losses[-1]:.5f}')
    ok &= same and lsame and step.replays == 3
    print(f'rank {rank}: {"OK" if ok else "MISMATCH"}', flush=True)
    step.close()                                               # graphs with NCCL kernels go before the process group
    del step, opt, tm
    torch.cuda.synchronize()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
