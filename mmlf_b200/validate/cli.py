"""Drop-in for ``python -m mmlf.validate.cli`` (/root/reference/mmlf/validate/cli.py:190-352): loads
``OUTPUT_DIR/checkpoint.pt``, rebuilds the model from its ``hyper_parameters``, reads the HCI4D scenes under DATASET
(``mmlf_b200.data.hci4d.HCI4D`` with the ``Shift(train_shift)`` transform, validate/cli.py:219), runs full-image inference
(BASE / UPR / DPP, or the ESE shift ensemble with ``--val_ensamble``; under torchrun the ESE members or the row bands of
one light field are sharded over the ranks), prints MSE / BadPix / the three KLD columns (computed on the GPU in float64,
``validate/metrics.py``) and dumps every scene through ``HCI4D.save_batch`` like the reference (:313).

A DATASET directory without scene sub-directories is an ERROR; ``--synthetic`` (extra flag) validates on seeded synthetic
scenes of ``--size`` pixels instead, with a warning, and writes no result files."""
import os
import sys
import time

import click
import numpy as np
import torch

from .. import parallel
from ..data import hci4d, synthetic
from ..model import loss
from ..model.ensamble import Ensamble
from ..model.feed_forward import FeedForward
from ..utils.dl import mpi_to_weights
from . import metrics


@click.command()
@click.argument('output_dir', type=click.Path(exists=True))
@click.argument('dataset', type=click.Path(exists=True))
@click.option('--model_invertible', is_flag=True, help='Use invertible architecture?')
@click.option('--model_discrete', is_flag=True, help='Discretize disparity output?')
@click.option('--val_loss_margin', default=15, help='Margin around each image to omit for the validation loss')
@click.option('--val_ensamble', is_flag=True, help='Use a network ensamble?')
@click.option('--val_disp_min', default=-3.5, help='Minimum disparity of dataset')
@click.option('--val_disp_max', default=3.5, help='Maximum disparity of dataset')
@click.option('--val_disp_step', default=0.1, help='Disparity increment for ensamble')
@click.option('--train_shift', default=0.0, type=float, help='Static shift to apply to off-center training datasets')
@click.option('--synthetic', 'synthetic_data', is_flag=True, help='[mmlf_b200] validate on synthetic scenes, not on DATASET')
@click.option('--size', default=128, help='[mmlf_b200] side length of the --synthetic validation scenes')
def main(output_dir, dataset, model_invertible, model_discrete, val_loss_margin, val_ensamble, val_disp_step,
         val_disp_min, val_disp_max, train_shift, synthetic_data, size):
    if model_invertible:
        raise NotImplementedError('INNs are not supported anymore')
    rank, world, local = parallel.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    state = torch.load(os.path.join(output_dir, 'checkpoint.pt'), map_location=dev)
    kwargs = state['hyper_parameters']
    kwargs.update({'model_discrete': model_discrete, 'val_disp_min': val_disp_min, 'val_disp_max': val_disp_max,
                   'train_shift': train_shift})
    shift = hci4d.Shift(float(train_shift))                      # applied even for 0.0, like validate/cli.py:219
    if synthetic_data:
        if rank == 0:
            print(f'WARNING: --synthetic: validating on SYNTHETIC light fields, not on {dataset!r}', file=sys.stderr)
        valset = synthetic.SyntheticLF(length=2, n=kwargs['model_views'], H=size, W=size, seed=3, name='val')
    else:
        if not any(f.is_dir() for f in os.scandir(dataset)):
            raise click.UsageError(f'DATASET {dataset!r} holds no scene directories (HCI4D layout, hci4d.py:82-95); '
                                   'pass --synthetic to validate on synthetic scenes instead')
        valset = hci4d.HCI4D(dataset, nviews=(kwargs['model_views'], kwargs['model_views']), transform=shift, device=dev)
    model = FeedForward(**kwargs).to(dev)
    mse_fn, bad_pix_fn = loss.MaskedMSELoss(), loss.MaskedBadPix()
    print('Loading model...')
    model.load_state_dict(state['model_state_dict'])
    if val_ensamble:
        model = Ensamble(model, val_disp_min, val_disp_max, val_disp_step)
    print('Number of parameters:', sum(p.numel() for p in model.parameters()))
    with torch.no_grad():
        model.eval()
        mse_avg = bad_pix_avg = kld_avg = kld_mm_avg = kld_um_avg = 0.0
        runtime = 0.0
        for i in range(len(valset)):
            print(f'Processing scene {i}...')
            t_start = time.time()
            item = valset[i]
            if synthetic_data:                         # the dataset transform of validate/cli.py:219, on the GPU: views,
                item = list(item)                      # gt and mpi[:, 4] (hci4d.py:984-988)
                item[:8] = [torch.as_tensor(t).to(dev) for t in item[:8]]
                item = shift(tuple(item))
            h_views, v_views, i_views, d_views, center, gt, mpi, _ = [torch.as_tensor(t).to(dev).unsqueeze(0) for t in item[:8]]
            index = np.atleast_2d(np.asarray(item[8]))
            mask = loss.create_mask_margin(gt.shape, val_loss_margin).to(dev)
            if world > 1 and not val_ensamble:
                # one light field, rows sharded over the ranks in bands with a `model_radius` halo
                output = parallel.banded_forward(model, [h_views, v_views, i_views, d_views],
                                                 radius=kwargs.get('model_radius', 11))
            else:
                output = model(h_views, v_views, i_views, d_views)
            mse = mse_fn(output, gt, mask)
            bad_pix = bad_pix_fn(output, gt, mask)
            mse_avg += mse.item()
            bad_pix_avg += bad_pix.item()
            mean_np = output['mean'].cpu().numpy()
            runtime = time.time() - t_start
            print(mse.item(), bad_pix.item())
            # distribution metrics on the GPU (validate/cli.py:286-325), float64 like the numpy originals
            mpi = mpi.float()
            dist_gt = mpi_to_weights(mpi, val_disp_min, val_disp_max, 108).double().contiguous()
            mm_mask = metrics.multimodal_mask(mpi)
            if val_ensamble:
                # the reference hands exp(logvars) to lmm_to_discrete, which exponentiates again (:304, :93): kept
                dist = metrics.lmm_to_discrete(108, val_disp_min, val_disp_max, output['means'], torch.exp(output['logvars']))
            elif model_discrete:
                dist = output['posterior'].double().contiguous()
            elif kwargs.get('model_uncert'):
                dist = metrics.laplace_to_discrete(108, val_disp_min, val_disp_max, output['mean'], output['logvar'])
            else:
                dist = metrics.mean_to_discrete(108, val_disp_min, val_disp_max, output['mean']).contiguous()
            if not synthetic_data and rank == 0:
                # result dump of validate/cli.py:287-314 (the only device -> host copies of the per-scene tensors)
                get = lambda k: None if output.get(k) is None else output[k].cpu().numpy()  # noqa: E731
                lmm = None
                if get('means') is not None and get('logvars') is not None:
                    lmm = np.stack([get('means'), np.exp(get('logvars'))], 0)
                valset.save_batch(output_dir, index, mean_np, get('logvar'), runtime, lmm, get('scores'), get('posterior'))
            if dist.shape[1] == 108:
                kld = metrics.kl_divergence(dist, dist_gt)               # three calls on the same, in-place normalised
                kld_mm = metrics.kl_divergence(dist, dist_gt, mm_mask)   # arrays, as validate/cli.py:323-325
                kld_um = metrics.kl_divergence(dist, dist_gt, 1.0 - mm_mask)
                print(kld_um, kld_mm, kld)
                kld_avg, kld_mm_avg, kld_um_avg = kld_avg + kld, kld_mm_avg + kld_mm, kld_um_avg + kld_um
        mse_avg /= (i + 1)
        bad_pix_avg /= (i + 1)
        kld_avg, kld_mm_avg, kld_um_avg = kld_avg / (i + 1), kld_mm_avg / (i + 1), kld_um_avg / (i + 1)
    if rank == 0:
        print('MSE & BadPix007 & KLD_UM & KLD_MM & KLD & - & TIME \\\\')
        print(f'{mse_avg:.3f} & {bad_pix_avg:.3f} & {kld_um_avg:.3f} & {kld_mm_avg:.3f} & {kld_avg:.3f} & - & {runtime:.3f} \\\\')
    return 0


if __name__ == '__main__':
    sys.exit(main())
