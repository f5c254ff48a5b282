"""Drop-in for ``python -m mmlf.train.cli`` (/root/reference/mmlf/train/cli.py): same click options and defaults, same
loop (warm start, cooling, periodic validation, log.csv, checkpoint.pt through ModelSaver).

Differences, all on the B200 side of the boundary:
  * one process per GPU under torchrun (NCCL all-reduce of one flat gradient bucket + the loss normalisers) replaces
    the single-process ``torch.nn.DataParallel`` of train/cli.py:159;
  * ``torch.optim.Adam`` is replaced by the fused flat-buffer Adam (same update rule, same state_dict format), and the
    body of the loop (train/cli.py:243-258: forward, loss, backward, optimizer step) is ``mmlf_b200.train.step.TrainStep``:
    one captured CUDA graph per batch shape, replayed every iteration;
  * ``--train_trainset`` / ``--train_valset`` are HCI4D directories exactly as upstream (train/cli.py:95-104).  The scenes
    are decoded once, cached in HBM, and the transform chain of train/cli.py:72-92 (static ``Shift``, then the random
    augmentations or, with ``--train_no_data_augment``, the plain crop) runs as the fused GPU gather of
    ``mmlf_b200.data.augment`` with its parameters drawn by the host ``random`` in the reference's order -- there are no
    CPU dataset workers, so ``--train_num_workers`` is accepted and unused on that path;
  * a dataset directory that does not exist is an ERROR.  ``--synthetic`` (extra flag) asks for the seeded synthetic
    light-field source instead (no HCI data offline); it prints a warning, and records ``synthetic: True`` in the
    checkpoint's hyper-parameters so that such a checkpoint cannot be mistaken for a real run;
  * every rank starts from rank 0's parameters and buffers (broadcast after construction / ``--train_resume``), like
    DataParallel's per-step replication from device 0; BatchNorm running statistics stay rank-local during training
    (DataParallel's replicas do the same) and rank 0's are the ones checkpointed.
``--max_iterations`` (extra option, default 0 = run forever like the reference) bounds the loop for tests.
"""
import os
import sys
import time

import click
import torch

from .. import parallel
from ..data import synthetic
from ..model import loss
from ..model.ensamble import Ensamble
from ..utils.dl import mpi_to_weights
from ..model.feed_forward import FeedForward
from ..optim import FusedAdam
from ..utils.dl import ModelSaver
from .step import TrainStep


# Option table of the reference CLI (train/cli.py:17-59): same names, defaults and types, applied programmatically;
# the last two entries are extensions of this package.
_OPTIONS = [
    ('--model_ksize', dict(default=2)),
    ('--model_in_blocks', dict(default=3)),
    ('--model_out_blocks', dict(default=8)),
    ('--model_chs', dict(default=70)),
    ('--model_views', dict(default=9)),
    ('--model_cross', dict(is_flag=True)),
    ('--model_uncert', dict(is_flag=True)),
    ('--model_discrete', dict(is_flag=True)),
    ('--model_unet', dict(is_flag=True)),
    ('--model_invertible', dict(is_flag=True)),
    ('--model_clamp', dict(default=0.7)),
    ('--model_act_norm', dict(default=0.7)),
    ('--model_act_norm_type', dict(default='SOFTPLUS')),
    ('--model_soft_permutation', dict(is_flag=True)),
    ('--model_no_batchnorm', dict(is_flag=True)),
    ('--model_batchnorm_momentum', dict(default=0.1)),
    ('--train_trainset', dict(default='../lf-dataset/additional')),
    ('--train_valset', dict(default='../lf-dataset/training')),
    ('--train_no_data_augment', dict(is_flag=True)),
    ('--train_num_workers', dict(default=4)),
    ('--train_lr', dict(default=1e-5)),
    ('--train_bs', dict(default=1)),
    ('--train_ps', dict(default=32)),
    ('--train_beta', dict(default=1.0)),
    ('--train_mae_threshold', dict(default=0.02)),
    ('--train_max_downscale', dict(default=4)),
    ('--train_resume', dict(is_flag=True)),
    ('--train_loss_padding', dict(default=None, type=float)),
    ('--train_shift', dict(default=0.0, type=float)),
    ('--train_loss_multimodal', dict(is_flag=True)),
    ('--train_loss_strongest', dict(is_flag=True)),
    ('--train_eval_mode', dict(is_flag=True)),
    ('--train_eval_mode_start', dict(default=0)),
    ('--train_warm_start', dict(is_flag=True)),
    ('--train_cooling', dict(default=0)),
    ('--val_interval', dict(default=100)),
    ('--val_loss_margin', dict(default=15)),
    ('--val_ensamble', dict(is_flag=True)),
    ('--val_disp_min', dict(default=-3.5)),
    ('--val_disp_max', dict(default=3.5)),
    ('--val_disp_step', dict(default=0.1)),
    ('--max_iterations', dict(default=0)),
    ('--gpu_augment', dict(is_flag=True)),
    (('--synthetic', 'synthetic_data'), dict(is_flag=True)),
]


def _with_options(fn):
    for name, kw in reversed(_OPTIONS):
        fn = click.option(*((name,) if isinstance(name, str) else name), **kw)(fn)
    return click.argument('output_dir', type=click.Path(exists=True))(fn)


@click.command()
@_with_options
def main(output_dir, max_iterations, gpu_augment, synthetic_data, **kwargs):
    assert not (kwargs['train_loss_strongest'] and kwargs['train_loss_multimodal'])
    if kwargs['model_invertible']:
        raise NotImplementedError('INNs are not supported anymore')          # train/cli.py:252
    kwargs['model_radius'] = (kwargs['model_in_blocks'] + kwargs['model_out_blocks']) * ((kwargs['model_ksize'] + 1) // 2)
    if kwargs['val_ensamble']:
        kwargs['model_uncert'] = True                                          # train/cli.py:68-69

    rank, world, local = parallel.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    ps = kwargs['train_ps']
    import random
    from ..data import hci4d
    from ..data.augment import GpuAugmenter
    random.seed(1234 + rank)
    sampler = None
    per_rank_bs = max(1, kwargs['train_bs'] // world)
    if synthetic_data:
        if rank == 0:
            print('WARNING: --synthetic: training and validating on SYNTHETIC light fields, not on '
                  f"{kwargs['train_trainset']!r} / {kwargs['train_valset']!r}", file=sys.stderr)
        kwargs['synthetic'] = True
        trainset = synthetic.SyntheticLF(length=4096, n=kwargs['model_views'], H=ps, W=ps, seed=1)
        sampler = torch.utils.data.distributed.DistributedSampler(trainset, world, rank, shuffle=True) if world > 1 else None
        trainloader = torch.utils.data.DataLoader(trainset, batch_size=per_rank_bs, shuffle=sampler is None,
                                                  sampler=sampler, num_workers=kwargs['train_num_workers'])
        valset = synthetic.SyntheticLF(length=2, n=kwargs['model_views'], H=4 * ps, W=4 * ps, seed=2, name='val')
        train_scenes = None
    else:
        for opt in ('train_trainset', 'train_valset'):
            if not os.path.isdir(kwargs[opt]):
                raise click.UsageError(f"--{opt} {kwargs[opt]!r} is not a directory (HCI4D layout, train/cli.py:95-104); "
                                       'pass --synthetic to train on the synthetic light-field source instead')
        nv = (kwargs['model_views'], kwargs['model_views'])
        train_scenes = hci4d.HCI4D(kwargs['train_trainset'], nviews=nv, cache=True, length=4096, device=dev)
        valset = hci4d.HCI4D(kwargs['train_valset'], nviews=nv, cache=True, device=dev)
        trainloader = None

    model = FeedForward(**kwargs).to(dev)
    optimizer = FusedAdam(model.parameters(), lr=kwargs['train_lr'])
    if kwargs['train_loss_multimodal']:
        loss_fn, loss_uncert_fn = loss.MultiMaskedL1Loss(), loss.ImprovedMultiUncertaintyL1Loss()
    else:
        loss_fn, loss_uncert_fn = loss.MaskedL1Loss(), loss.ImprovedUncertaintyL1Loss()
    mse_fn, bad_pix_fn = loss.MaskedMSELoss(), loss.MaskedBadPix()

    augmenter = None
    if train_scenes is not None or (gpu_augment and not kwargs['train_no_data_augment']):
        # the reference's transform chain (train/cli.py:72-92) on the GPU: scenes resident in HBM, parameters drawn by the
        # host `random` in the reference's order, one gather kernel per batch (mmlf_b200/data/augment.py)
        if train_scenes is not None:
            scenes = train_scenes.data
        else:
            side = max(4 * ps, kwargs['train_max_downscale'] * (ps + 17))
            scenes = [synthetic.SyntheticLF(length=8, n=kwargs['model_views'], H=side, W=side, seed=7)[j] for j in range(8)]
        augmenter = GpuAugmenter(scenes, dev)
        if kwargs['train_shift'] != 0.0:                                       # static Shift first (train/cli.py:89-90)
            for j in range(augmenter.S):
                st = [augmenter.stacks[j, k] for k in range(4)]
                hci4d.Shift(float(kwargs['train_shift']))(tuple(st) + (None, augmenter.gt[j], augmenter.mpi[j]))
        plain = bool(kwargs['train_no_data_augment'])

        def gpu_batches():
            while True:
                ids, params = augmenter.draw(per_rank_bs, ps, random, kwargs['train_max_downscale'], plain=plain)
                yield augmenter(ids, params, plain=plain)
        trainloader = gpu_batches()
    static_shift = float(kwargs['train_shift']) if (augmenter is None and kwargs['train_shift'] != 0.0) else None

    i = 0
    if kwargs['train_resume']:                                                # train/cli.py:137-157
        print('Resume training...')
        state = torch.load(os.path.join(output_dir, 'checkpoint.pt'), map_location=dev)
        for k in list(state['model_state_dict']):
            if 'tmp' in k:
                del state['model_state_dict'][k]
        model.load_state_dict(state['model_state_dict'])
        optimizer.load_state_dict(state['optimizer_state_dict'])
        for param_group in optimizer.param_groups:
            param_group['lr'] = kwargs['train_lr']
        i = state['iteration']
    # one replica: every rank continues from rank 0's parameters, buffers and Adam moments (DataParallel replicates from
    # device 0 every step, train/cli.py:159; here the ranks stay identical because they apply the same summed gradient)
    parallel.broadcast_module_(model)
    optimizer.broadcast_state_()
    val_model = Ensamble(model, **kwargs) if kwargs['val_ensamble'] else model
    if kwargs['model_uncert']:
        step_loss = 'multi_upr' if kwargs['train_loss_multimodal'] else 'upr'
    elif kwargs['model_discrete']:
        step_loss = 'ce'
    else:
        step_loss = 'multi_l1' if kwargs['train_loss_multimodal'] else 'l1'
    train_step = TrainStep(model, optimizer, step_loss,
                           ce_from_gt=kwargs['model_discrete'] and not kwargs['train_loss_multimodal'])

    log = None
    header = f'{"iter":>7}, loss_train,   loss_val,        mse, badpix_007, time_elapsed'
    if rank == 0:
        log = open(os.path.join(output_dir, 'log.csv'), 'a' if kwargs['train_resume'] else 'w')
        print(header)
        if not kwargs['train_resume']:
            print(header, file=log)
    model_saver = ModelSaver(only_best=False)
    loss_val_avg = mse_avg = bad_pix_avg = 0.0
    dims = (2 if kwargs['model_cross'] else 4) * kwargs['model_views'] * 3
    time_start = 0
    epoch = 0
    margin_masks = {}

    def val_batches():
        """The validation scenes one by one with a batch dimension of 1 (DataLoader(valset, batch_size=1), train/cli.py:101)."""
        for j in range(len(valset)):
            item = valset[j]
            yield [torch.as_tensor(t).unsqueeze(0) for t in item]

    # The log line of iteration i (train/cli.py:296-299) needs the loss on the host.  Reading it right after the step is
    # enqueued would stall the host until the step has finished and leave the GPU idle while the next batch is prepared
    # (GPU augmentation: ~2 ms of host work per 64 patches), so the read-back is deferred until the next batch has been
    # prepared and its kernels are queued behind the running step -- and always happens before the next step overwrites
    # the loss buffer.  Same lines, same order.
    pending = []

    def flush_log():
        while pending:
            it_, loss_dev, lv, ms_, bp, te = pending.pop(0)
            line = f'{it_:>7}, {loss_dev.item():.8f}, {lv:.8f}, {ms_:.8f}, {bp:.8f}, {te:.8f}'
            print(line)
            print(line, file=log, flush=True)

    while True:
        if sampler is not None:
            sampler.set_epoch(epoch)                                          # reshuffle every pass, like shuffle=True
        epoch += 1
        for data in trainloader:
            flush_log()                                   # iteration i - 1, now that batch i is on its way
            h_views, v_views, i_views, d_views, center, gt, mpi, mask, index = data
            if kwargs['train_loss_strongest']:
                inds = torch.max(mpi[:, :, 3, :, :], dim=1)[1].unsqueeze(1)
                gt = torch.gather(mpi[:, :, 4, :, :], dim=1, index=inds).squeeze()
            mkey = (tuple(mask.shape), str(mask.device))
            if mkey not in margin_masks:
                margin_masks[mkey] = loss.create_mask_margin(mask.shape, 11).to(mask.device)
            mask = mask.int() * margin_masks[mkey]                               # train/cli.py:194
            h_views, v_views, i_views, d_views = (t.to(dev, non_blocking=True) for t in (h_views, v_views, i_views, d_views))
            gt, mpi, mask = gt.to(dev), mpi.to(dev).float(), mask.to(dev)
            if static_shift is not None:        # Shift(train_shift) of train/cli.py:89-90 on the DataLoader path, on the GPU
                hci4d.Shift(static_shift)((h_views, v_views, i_views, d_views))
                gt = gt - static_shift
                mpi[:, :, 4] -= static_shift
            mask_padding = None
            if kwargs['train_loss_padding'] is not None:
                if kwargs['train_loss_multimodal']:
                    mpi[:, :, 3, :, :] *= (torch.abs(mpi[:, :, 4, :, :]) < kwargs['train_loss_padding']).float()
                else:
                    mask_padding = (torch.abs(gt) < kwargs['train_loss_padding']).int()
            if kwargs['train_loss_multimodal']:
                gt = mpi
            if kwargs['train_eval_mode'] and i >= kwargs['train_eval_mode_start']:   # train/cli.py:227-230
                model.eval()
            else:
                model.train()
            if kwargs['train_warm_start'] and i <= 1000:                      # train/cli.py:233-236
                for g in optimizer.param_groups:
                    g['lr'] = kwargs['train_lr'] * float(i) / 1000.0
            if kwargs['train_cooling'] > 0 and i >= kwargs['train_cooling']:
                lr = kwargs['train_lr'] / (10.0 ** (i / kwargs['train_cooling'] - 1.0))
                for g in optimizer.param_groups:
                    g['lr'] = lr
            # forward + loss + backward + gradient all-reduce + Adam: one graph replay (train/cli.py:243-258).  For
            # --model_discrete the class targets (utils/dl.py:109-157) are built inside the loss kernel from gt, or on the
            # GPU from the MPI planes for --train_loss_multimodal.
            if kwargs['model_discrete'] and kwargs['train_loss_multimodal']:
                target = mpi_to_weights(mpi, kwargs['val_disp_min'], kwargs['val_disp_max'], dims)
            else:
                target = gt
            loss_train = train_step(h_views, v_views, i_views, d_views, target, mask, mask_padding)
            time_elap = time.time() - time_start

            if i % kwargs['val_interval'] == 0:
                with torch.no_grad():
                    model.eval()
                    loss_val_avg = mse_avg = bad_pix_avg = 0.0
                    for j, vdata in enumerate(val_batches()):
                        vh, vv, vi, vd, center, vgt, vmpi, _, index = vdata
                        vh, vv, vi, vd = (t.to(dev) for t in (vh, vv, vi, vd))
                        vgt, vmpi = vgt.to(dev), vmpi.to(dev).float()
                        vmask = loss.create_mask_margin(vgt.shape, kwargs['val_loss_margin']).to(dev)
                        output = val_model(vh, vv, vi, vd)
                        tgt = vmpi if kwargs['train_loss_multimodal'] else vgt
                        loss_val = (loss_uncert_fn if kwargs['model_uncert'] else loss_fn)(output, tgt, vmask)
                        loss_val_avg += loss_val.item()
                        mse_avg += mse_fn(output, vgt, vmask).item()
                        bad_pix_avg += bad_pix_fn(output, vgt, vmask).item()
                    j += 1
                    loss_val_avg, mse_avg, bad_pix_avg = loss_val_avg / j, mse_avg / j, bad_pix_avg / j
                    if rank == 0:
                        model_saver(os.path.join(output_dir, 'checkpoint.pt'), model, optimizer, kwargs, None, i,
                                    loss_val_avg)
            if rank == 0:
                pending.append((i, loss_train, loss_val_avg, mse_avg, bad_pix_avg, time_elap))
            i += 1
            time_start = time.time()
            if max_iterations and i >= max_iterations:
                flush_log()
                train_step.close()                    # captured NCCL kernels must go before the process group does
                return 0


if __name__ == '__main__':
    sys.exit(main())
