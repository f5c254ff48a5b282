"""Drop-in for ``python -m mmlf.validate.cli`` (/root/reference/mmlf/validate/cli.py:190-352), model side: loads
``OUTPUT_DIR/checkpoint.pt``, rebuilds the model from its ``hyper_parameters``, runs full-image inference (BASE / UPR /
DPP, or the ESE shift ensemble with ``--val_ensamble``, members sharded over the ranks under torchrun) and reports MSE /
BadPix.  The numpy KLD / NLL post-processing helpers and the PFM/PNG result dump are CPU-side and out of scope (SURVEY.md
section 2 row 8); DATASET is accepted for interface parity and replaced by synthetic scenes."""
import os
import sys
import time

import click
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
@click.option('--size', default=128, help='[mmlf_b200] side length of the synthetic validation scenes')
def main(output_dir, dataset, model_invertible, model_discrete, val_loss_margin, val_ensamble, val_disp_step,
         val_disp_min, val_disp_max, train_shift, size):
    rank, world, local = parallel.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    state = torch.load(os.path.join(output_dir, 'checkpoint.pt'), map_location=dev)
    kwargs = state['hyper_parameters']
    kwargs.update({'model_discrete': model_discrete, 'val_disp_min': val_disp_min, 'val_disp_max': val_disp_max,
                   'train_shift': train_shift})
    valset = synthetic.SyntheticLF(length=2, n=kwargs['model_views'], H=size, W=size, seed=3, name='val')
    valloader = torch.utils.data.DataLoader(valset, batch_size=1, shuffle=False, num_workers=1)
    model = FeedForward(**kwargs).to(dev)
    mse_fn, bad_pix_fn = loss.MaskedMSELoss(), loss.MaskedBadPix()
    print('Loading model...')
    model.load_state_dict(state['model_state_dict'])
    if val_ensamble:
        model = Ensamble(model, val_disp_min, val_disp_max, val_disp_step)
    print('Number of parameters:', sum(p.numel() for p in model.parameters()))
    shift = hci4d.Shift(float(train_shift)) if train_shift != 0.0 else None
    with torch.no_grad():
        model.eval()
        mse_avg = bad_pix_avg = kld_avg = kld_mm_avg = kld_um_avg = 0.0
        runtime = 0.0
        for i, data in enumerate(valloader):
            print(f'Processing scene {i}...')
            t_start = time.time()
            data = [t.to(dev) if isinstance(t, torch.Tensor) else t for t in data]
            if shift is not None:                      # the dataset transform of validate/cli.py:219 runs on the GPU
                data[:4] = shift(tuple(data[:4]))[:4]
                data[5] = data[5] - float(train_shift)
            h_views, v_views, i_views, d_views, center, gt, mpi, _, index = data
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
            _ = output['mean'].cpu()
            runtime = time.time() - t_start
            print(mse.item(), bad_pix.item())
            # distribution metrics on the GPU (validate/cli.py:286-325), float64 like the numpy originals
            mpi = mpi.to(dev).float()
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
