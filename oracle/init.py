"""Oracle (test infrastructure): the reference's default parameter initialisation without the reference and without the
product package -- the same ``torch.nn`` containers, created in the same order as
/root/reference/mmlf/model/feed_forward.py:104-187 (``block`` :106-137, ``init_in_net`` :139-160, ``init_out_net``
:162-187), so that ``torch.manual_seed(s)`` yields the reference's initial ``state_dict`` bit for bit
(tests/test_host_logic.py::test_init_matches_reference_rng_stream pins the product's twin of this against a reference
fixture; tests/test_oracle_golden.py pins this one against the product's).  Used by bench.py's CPU arms."""
import numpy as np


def default_state(model_chs=70, model_views=9, model_cross=False, model_uncert=False, model_discrete=False,
                  model_in_blocks=3, model_out_blocks=8, model_no_batchnorm=False, model_batchnorm_momentum=0.1, seed=0,
                  **_):
    import torch
    import torch.nn as nn

    def block(ch_in, ch_out=None, out_bn_relu=True):
        ch_out = ch_in if ch_out is None else ch_out
        layers = [nn.Conv2d(ch_in, ch_out, 2, padding=1), nn.ReLU(), nn.Conv2d(ch_out, ch_out, 2, padding=0)]
        if out_bn_relu:
            if not model_no_batchnorm:
                layers.append(nn.BatchNorm2d(ch_out, momentum=model_batchnorm_momentum))
            layers.append(nn.ReLU())
        return nn.Sequential(*layers)

    def in_net():
        return nn.Sequential(block(model_views * 3, model_chs), *[block(model_chs) for _ in range(model_in_blocks - 1)])

    torch.manual_seed(seed)
    net = nn.Module()
    net.in_net_hv = in_net()
    if not model_cross:
        net.in_net_id = in_net()
    width = (2 if model_cross else 4) * model_chs
    steps = (2 if model_cross else 4) * model_views * 3
    oc = 2 if model_uncert else (steps if model_discrete else 1)
    net.out_net = nn.Sequential(*[block(width) for _ in range(model_out_blocks - 1)], block(width, oc, False))
    return {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
