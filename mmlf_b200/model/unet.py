"""Parameter containers of the ``--model_unet`` out-net (/root/reference/mmlf/model/unet.py:8-132 as instantiated by
feed_forward.py:189-204: ``UNet(chs, out_chs, depth=5, padding=True, batch_norm=True)``, 'upconv' up-sampling).

Containers only: the module tree fixes the ``state_dict`` keys (``down_path.{i}.block.{0,2,3,5}``, ``up_path.{i}.up``,
``up_path.{i}.conv_block.block.*``, ``last``) and the order in which ``torch.manual_seed`` initialises them; execution is
``mmlf_b200.engine_generic.GenericEngine`` on the float32 layer kernels of csrc/generic.cu."""
import torch.nn as nn


def conv_block(cin, cout):
    """unet.py:80-101: [conv3x3 p1, ReLU, BatchNorm] x 2 -- note BatchNorm AFTER the ReLU, default momentum 0.1."""
    holder = nn.Module()
    holder.block = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(), nn.BatchNorm2d(cout),
                                 nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.ReLU(), nn.BatchNorm2d(cout))
    return holder


def up_block(cin, cout):
    """unet.py:104-131: ConvTranspose2d(k 2, stride 2), then the conv block on cat([up, centre-cropped bridge])."""
    holder = nn.Module()
    holder.up = nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2)
    holder.conv_block = conv_block(cin, cout)
    return holder


class UNet(nn.Module):
    def __init__(self, in_channels, n_classes, depth=5, wf=6):
        super().__init__()
        self.depth = depth
        widths = [2 ** (wf + i) for i in range(depth)]
        self.down_path = nn.ModuleList()
        prev = in_channels
        for wdt in widths:
            self.down_path.append(conv_block(prev, wdt))
            prev = wdt
        self.up_path = nn.ModuleList()
        for wdt in reversed(widths[:-1]):
            self.up_path.append(up_block(prev, wdt))
            prev = wdt
        self.last = nn.Conv2d(prev, n_classes, kernel_size=1)

    def forward(self, x):
        raise RuntimeError('mmlf_b200.model.unet.UNet is a parameter container; run it through FeedForward')
