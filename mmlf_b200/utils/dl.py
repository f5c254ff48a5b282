"""Drop-in for the hot-path part of ``mmlf.utils.dl`` (/root/reference/mmlf/utils/dl.py): the checkpoint writer and
the disparity <-> bin helpers, plus ``save_img`` (PNG writer used by ``HCI4D.save_batch``).  ``BatchIter`` is an unused
helper upstream and not provided."""
import numpy as np
import torch

from .. import ops


class ModelSaver:
    """utils/dl.py:7-74: writes the same ``checkpoint.pt`` dictionary (model_state_dict without ``module.`` prefix,
    optimizer_state_dict, hyper_parameters, epoch, iteration, loss)."""

    def __init__(self, only_best=False):
        self.only_best = only_best
        self.best_loss = None

    def __call__(self, fname, model, optimizer=None, hyper_parameters=None, epoch=None, iteraration=None, loss=None,
                 **kwargs):
        if self.only_best and loss is not None:
            if self.best_loss is not None and self.best_loss < loss:
                return
            self.best_loss = loss
        try:
            model_state_dict = model.module.state_dict()
        except AttributeError:
            model_state_dict = model.state_dict()
        optimizer_state_dict = None
        if optimizer is not None:
            optimizer_state_dict = optimizer.state_dict()
        state = {'model_state_dict': model_state_dict, 'optimizer_state_dict': optimizer_state_dict,
                 'hyper_parameters': hyper_parameters, 'epoch': epoch, 'iteration': iteraration, 'loss': loss}
        state.update(kwargs)
        torch.save(state, fname)


def _bins(start, stop, n_steps, device):
    return ops.torch_bins(start, stop, n_steps, device)


def reg_to_class(arr, start, stop, n_steps):
    """utils/dl.py:109-131 on the GPU: (B, H, W) -> (B, n_steps, H, W) one-hot float32."""
    step = (stop - start) / n_steps
    return ops.reg_to_class_op(arr.to(torch.float32).contiguous(), _bins(start, stop, n_steps, arr.device), step / 2.0)


def mpi_to_weights(arr, start, stop, n_steps):
    """utils/dl.py:134-157 on the GPU: (B, K, 5, H, W) -> (B, n_steps, H, W)."""
    step = (stop - start) / n_steps
    return ops.mpi_to_weights_op(arr.to(torch.float32).contiguous(), _bins(start, stop, n_steps, arr.device),
                                 step / 2.0)


def class_to_reg(arr, start, stop, n_steps):
    """utils/dl.py:160-182 (small reduction over the bin axis; plain torch, not on the per-step path because the
    DPP head kernel computes `mean` directly)."""
    result = torch.linspace(start, stop, n_steps).view((1, -1, 1, 1)).to(arr.device)
    return torch.sum(result * arr, 1)


def save_img(fname, arr):
    """utils/dl.py:75-105: (3, h, w) rgb or (h, w) grey array / tensor -> 8-bit PNG; values outside [0, 1] are min-max
    normalised first.  The float -> uint8 step is skimage.img_as_ubyte's: ``rint(x * 255)``.  PIL instead of skimage
    (not installed in this image); host-side file I/O."""
    from PIL import Image
    if not isinstance(arr, np.ndarray):
        arr = arr.detach().cpu().numpy()
    arr = np.asarray(arr, dtype=np.float64 if arr.dtype == np.float64 else np.float32)
    a_min, a_max = np.min(arr), np.max(arr)
    if a_min < 0.0 or a_max > 1.0:
        arr = (arr - a_min) / (a_max - a_min)
    if arr.ndim == 3:
        arr = np.transpose(arr, (1, 2, 0))
    img = np.clip(np.rint(arr * 255.0), 0, 255).astype(np.uint8)
    Image.fromarray(img).save(fname)
